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
sys.path.insert(0, ROOT)
from bench import layer_bytes  # noqa: E402  (one model of the dataflow for the bench line and this table)
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
    by = layer_bytes(r, names, B, T, rates, 2)
    if r["kind"] < 0:
        fl = 0; shape = "mel [16,80,800]"; path = "mel_to_operand"
    else:
        cin, cout, k = r["c_in"], r["c_out"], r["k"]
        if n == "conv_pre":
            fl = 2 * B * T * cin * cout * k
        elif n.startswith("ups."):
            i = int(n.split(".")[1]); fl = 2 * B * T * math.prod(rates[:i]) * cin * cout * k
        elif n == "conv_post":
            fl = 2 * B * T * 256 * cin * k
        else:
            i = int(n.split(".")[1]) // 3
            fl = 2 * B * T * math.prod(rates[:i + 1]) * cout * cin * k
            if ".convs2." in n and n.replace(".convs2.", ".convs1.") not in names:
                fl *= 2
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
