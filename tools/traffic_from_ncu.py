#!/usr/bin/env python
"""profiles/traffic_<precision>.json from an ncu launch list of ONE forward (tools/ncu_one_forward.py under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
--launch-skip 61 --launch-count 61 --csv`): DRAM bytes per step, summed over the tcgen05 launches and over all.

    python tools/traffic_from_ncu.py profiles/r2_v10_ncu_launches_bf16.csv bf16 [in_situ.csv]

The optional second list is the same pass with `--cache-control none` (no L2 flush between kernels): what the step
really moves, including what one launch leaves in L2 for the next (profiles/r2_tile_order_l2.md).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def totals(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = {}
    for r in rows[1:]:
        if r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            k = per.setdefault(r[ii], {"kernel": r[ki], "bytes": 0.0})
            k["bytes"] += float(r[vi].replace(",", "")) * UNIT[r[ui]]
    tc = [v for v in per.values() if any(s in v["kernel"] for s in ("conv_tc", "conv_pair", "conv_chain"))]
    return {"tcgen05_kernels": {"launches": len(tc), "dram_bytes_per_step": sum(v["bytes"] for v in tc)},
            "all_kernels": {"launches": len(per), "dram_bytes_per_step": sum(v["bytes"] for v in per.values())}}


def main():
    src, prec = sys.argv[1], sys.argv[2]
    out = {"source": f"{os.path.relpath(src, ROOT)} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                     "--clock-control none, the second forward of tools/ncu_one_forward.py = one step of the bench workload; "
                     "ncu flushes L2 before every kernel)",
           "workload": "HiFi-GAN V1 16 x 800 frames, " + prec}
    out.update(totals(src))
    if len(sys.argv) > 3:
        out["in_situ"] = dict(source=f"{os.path.relpath(sys.argv[3], ROOT)} (same pass with --cache-control none)", **totals(sys.argv[3]))
    json.dump(out, open(os.path.join(ROOT, "profiles", f"traffic_{prec}.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
