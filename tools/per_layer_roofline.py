#!/usr/bin/env python
"""Per-launch roofline table from a bench_layers_*.json (bench.py / hg_profile_forward):

    python tools/per_layer_roofline.py profiles/r1_v8_layers_bf16.json > profiles/r1_per_layer_roofline_bf16.md

For every launch: algorithmic FLOPs (2 x MACs, all taps), algorithmic HBM bytes of
this implementation's dataflow (bf16 operand copies 2 B, fp32 residual stream 4 B; unique elements
read + written, halo re-reads and weights excluded), the roofline time
max(FLOPs / tensor peak, bytes / HBM peak) with the measured peaks of MEASURED_PEAKS.json, and the
fraction of it the measured CUDA-event time reaches."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = json.load(open(sys.argv[1]))
B, T = 16, 800
rates = [8, 8, 2, 2]
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
TF = float(peaks.get("bf16_tflops_sustained", 1402.0)) * 1e12
BW = float(peaks.get("hbm_gbs", 6555.8)) * 1e9
names = {r["name"] for r in rows}
print("| launch | shape | kernel path | ms | GFLOP | MB | bound | roofline ms | % of roofline | TFLOP/s | GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
tot_ms = tot_roof = 0.0
for r in rows:
    n = r["name"]
    if r["kind"] < 0:
        by = B * T * (80 * 4 + 128 * 2); fl = 0; shape = "mel [16,80,800]"; path = "mel_to_operand"
    else:
        cin, cout, k = r["c_in"], r["c_out"], r["k"]
        if n == "conv_pre":
            L = T; fl = 2 * B * L * cin * cout * k; by = B * L * (128 * 2 + cout * 2)
        elif n.startswith("ups."):
            i = int(n.split(".")[1]); Lin = T * math.prod(rates[:i]); Lout = Lin * r["stride"]
            fl = 2 * B * Lin * cin * cout * k; by = B * (Lin * cin * 2 + Lout * cout * 6)
        elif n == "conv_post":
            L = T * 256; fl = 2 * B * L * cin * k; by = B * L * (cin * 4 + 4)
        else:
            blk = int(n.split(".")[1]); i = blk // 3; L = T * math.prod(rates[:i + 1]); e = B * L * cout
            fl = 2 * e * cin * k
            pair = ".convs2." in n and n.replace(".convs2.", ".convs1.") not in names
            last = n.endswith(".2")  # last pair of the block: MRF combine fused
            if ".convs1." in n:
                by = e * (2 + 2)
            else:
                by = e * (2 + 4)                      # operand in + residual in
                if pair:
                    fl *= 2
                if not last:
                    by += e * (4 + 2)                 # x out + a out
                else:
                    j = blk % 3
                    by += e * (4 if j > 0 else 0)     # xs in
                    by += e * (4 if (j < 2 or i == 3) else 2)  # xs / x out, or next-stage operand out
        shape = f"{cin}->{cout} k{k}" + (f" d{r['dilation']}" if r["kind"] == 0 else f" s{r['stride']}")
        path = r.get("kernel") or ("conv_pair_tc (fused pair)" if (".convs2." in n and n.replace(".convs2.", ".convs1.") not in names) else (
            "tcgen05" if r.get("tensor_core") else "cuda-core"))
    t_f, t_b = fl / TF * 1e3, by / BW * 1e3
    roof = max(t_f, t_b)
    tot_ms += r["ms"]; tot_roof += roof
    print(f"| {n} | {shape} | {path} | {r['ms']:.3f} | {fl / 1e9:.1f} | {by / 1e6:.0f} | {'tensor' if t_f >= t_b else 'hbm'} | {roof:.3f} | "
          f"{100 * roof / r['ms']:.0f} | {fl / r['ms'] / 1e9:.0f} | {by / r['ms'] / 1e6:.0f} |")
print(f"\nTotal measured {tot_ms:.2f} ms; sum of per-layer roofline times {tot_roof:.2f} ms -> {100 * tot_roof / tot_ms:.0f} % of the per-layer roofline "
      f"(tensor peak {TF / 1e12:.0f} TFLOP/s sustained, HBM {BW / 1e9:.0f} GB/s, both measured).")
