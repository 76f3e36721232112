#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` exports (profiles/r1_*_ncu_full_*.csv) into one table:

    python tools/ncu_full_summary.py profiles/r1_v10_ncu_full_tc2.csv profiles/r1_v10_ncu_full_pair.csv > profiles/r1_v10_ncu_full_summary.md

Per captured launch: duration, DRAM traffic and throughput, tensor-pipe activity, and the occupancy of
the L1 / shared-memory data pipe split by client — tensor-core operand reads (tc), LSU (shared +
global) and TMA fills — which is the resource the 64/32-channel layers run out of.
The TMA share is an estimate (fill bytes / 128 B per wavefront / elapsed cycles), so the three can
add up to slightly more than 100 % on a saturated pipe.
"""
import csv
import sys

COLS = [
    ("us", "gpu__time_duration.sum"),
    ("dram rd MB", "dram__bytes_read.sum"),
    ("dram wr MB", "dram__bytes_write.sum"),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("L1 pipe: tc %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("L1 pipe: lsu %", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("tma fill GB", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum"),
    ("smem bank conflicts M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

print("| kernel | " + " | ".join(c for c, _ in COLS) + " | L1 pipe total % (tc + lsu + tma) |")
print("|---|" + "---|" * (len(COLS) + 1))
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    H, U, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(H)}
    for r in data:
        name = r[ix["Kernel Name"]].replace("void ", "").replace("hg::", "").split("(")[0]
        vals, raw = [], {}
        for label, m in COLS:
            if m not in ix:
                vals.append("-")
                continue
            v = float(r[ix[m]].replace(",", "") or 0)
            u = U[ix[m]]
            if u in SCALE:
                v *= SCALE[u]
            raw[label] = v
            if "MB" in label:
                vals.append(f"{v / 1e6:.0f}")
            elif "GB" in label:
                vals.append(f"{v / 1e9:.2f}")
            elif "KB" in label:
                vals.append(f"{v / 1e3:.0f}")
            elif label.endswith(" M"):
                vals.append(f"{v / 1e6:.2f}")
            elif label in ("regs", "grid"):
                vals.append(f"{v:.0f}")
            else:
                vals.append(f"{v:.1f}")
        # TMA fills as a share of the data pipe: bytes / 128 B per wavefront / (cycles * SMs)
        cyc = float(r[ix["sm__cycles_elapsed.max"]].replace(",", "")) if "sm__cycles_elapsed.max" in ix else 0
        tma_pct = raw.get("tma fill GB", 0) / 128 / (cyc * 148) * 100 if cyc else 0
        tot = raw.get("L1 pipe: tc %", 0) + raw.get("L1 pipe: lsu %", 0) + tma_pct
        print(f"| {name} | " + " | ".join(vals) + f" | {tot:.0f} (tma {tma_pct:.0f}) |")
