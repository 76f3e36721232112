#!/usr/bin/env python
"""Per-launch device times of one forward of any generator config (CUDA events around every launch):

    python tools/profile_config.py v2_narrow 16 800 bf16
    python tools/profile_config.py v1 1 256 bf16          # cfg-1
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

CFG = {"v1": fx.V1, "v2_narrow": fx.V2_NARROW, "v3_rb2": fx.V3_RB2}
name, B, T, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
m = make_generator(CFG[name], precision=prec).cuda()
mel = fx.synthetic_mel(B, T, seed=7).cuda()
with torch.no_grad():
    for _ in range(3):
        m(mel)
    rows = m.profile_layers(mel)
    rows = m.profile_layers(mel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m(mel)
    e1.record()
    torch.cuda.synchronize()
print(f"# {name} {B} x {T} {prec}: {e0.elapsed_time(e1) / 20:.3f} ms per forward, {len(rows)} launches, sum of launches {sum(r['ms'] for r in rows):.3f} ms")
for r in rows:
    print(f"{r['name']:24s} {r.get('c_in', ''):>4} {r.get('c_out', ''):>4} k{r.get('k', '')} d{r.get('dilation', '')} {r['kernel']:22s} {r['ms']:.4f}")
if "--json" in sys.argv:
    print(json.dumps(rows))
