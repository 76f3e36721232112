#!/usr/bin/env python
"""Per-launch times of a small forward (BASELINE cfg-1: 1 x 256 frames) — the latency regime.
    python tools/latency_profile.py [B] [T] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
m = make_generator(fx.V1, precision=prec).cuda()
mel = fx.synthetic_mel(B, T, seed=7).cuda()
with torch.no_grad():
    for _ in range(3):
        m(mel)
    best = None
    for _ in range(5):
        rows = m.profile_layers(mel)
        if best is None:
            best = rows
        else:
            for a, b in zip(best, rows):
                a["ms"] = min(a["ms"], b["ms"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        m(mel)
    e1.record()
    torch.cuda.synchronize()
print(f"B={B} T={T} {prec}: forward {e0.elapsed_time(e1) / 50:.3f} ms; sum of per-launch minima {sum(r['ms'] for r in best):.3f} ms")
stage = {}
for r in best:
    n = r["name"]
    key = n if not n.startswith("resblocks.") else f"stage{int(n.split('.')[1]) // 3} k={r['k']} {'c1' if 'convs1' in n else 'c2/pair'}"
    s = stage.setdefault(key, [0.0, 0])
    s[0] += r["ms"]; s[1] += 1
for k, (ms, n) in stage.items():
    print(f"  {k:24s} n={n} {ms * 1e3:8.1f} us  ({ms / n * 1e3:6.1f} us each)")
