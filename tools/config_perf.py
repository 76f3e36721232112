#!/usr/bin/env python
"""Throughput of the other BASELINE configs on one GPU: cfg-1 (1 x 256 frames, latency regime),
cfg-4 (V2-style narrow, uic 128, 16 x 800) and V3-style ResBlock2, next to V1 16 x 800."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

SMALL = os.environ.get("HG_SMALL") == "1"
cases = [("cfg1_v1_1x256", fx.V1, 1, 256), ("cfg2_v1_16x800", fx.V1, 16, 800), ("cfg4_v2narrow_16x800", fx.V2_NARROW, 16, 800),
         ("v3_rb2_16x800", fx.V3_RB2, 16, 800)]
if SMALL:
    cases = [("v1_2x40", fx.V1, 2, 40), ("v2narrow_2x40", fx.V2_NARROW, 2, 40), ("v3_rb2_2x40", fx.V3_RB2, 2, 40)]
for name, cfg, B, T in cases:
    for prec in ("bf16", "fp32"):
        m = make_generator(cfg, precision=prec).cuda()
        mel = fx.synthetic_mel(B, T, seed=7).cuda()
        hop = m.hop_length
        with torch.no_grad():
            for _ in range(1 if SMALL else 3):
                y = m(mel)
            torch.cuda.synchronize()
            if SMALL:
                print(json.dumps({"config": name, "precision": prec, "ok": bool(torch.isfinite(y).all())}))
                continue
            reps = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                y = m(mel)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rec = {"config": name, "precision": prec, "ms": ms, "audio_sec_per_sec": B * T * hop / 22050 / (ms * 1e-3),
               "launches": m.kernel_launches(B, T)}
        if B * T <= 1024:  # latency regime: also time the CUDA-graph replay
            run = m.make_graphed(B, T)
            for _ in range(3):
                run(mel)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                run(mel)
            e1.record()
            torch.cuda.synchronize()
            rec["ms_cuda_graph"] = e0.elapsed_time(e1) / reps
            rec["audio_sec_per_sec_cuda_graph"] = B * T * hop / 22050 / (rec["ms_cuda_graph"] * 1e-3)
        print(json.dumps(rec))
