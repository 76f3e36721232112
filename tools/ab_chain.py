#!/usr/bin/env python
"""A/B of the fused-ResBlock kernel (HG_CHAIN=1, conv_chain_tc.cu) against the three fused pairs, interleaved in one process."""
import os, sys, hashlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from oracle import fixtures as fx
from _util import make_generator
mel = fx.synthetic_mel(16, 800, seed=7).cuda()
models = {}
for name, env in (("pairs", {"HG_CHAIN": "0"}), ("chain", {"HG_CHAIN": "1"})):
    os.environ.update(env)
    m = make_generator(fx.V1, precision="bf16").cuda()
    with torch.no_grad(): y = m(mel)
    torch.cuda.synchronize()
    for k in env: os.environ.pop(k)
    models[name] = (m, hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16])
res = {n: [] for n in models}
with torch.no_grad():
    for r in range(6):
        for n, (m, _) in models.items():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): m(mel)
            e1.record(); torch.cuda.synchronize()
            res[n].append(e0.elapsed_time(e1) / 10)
    for n, (m, d) in models.items():
        rows = m.profile_layers(mel); rows = m.profile_layers(mel)
        print(n, d, "forward ms mean %.3f min %.3f" % (sum(res[n]) / len(res[n]), min(res[n])), "launches", len(rows))
        for r in rows:
            if r["name"] in ("resblocks.6.convs2.0", "resblocks.6.convs2.1", "resblocks.6.convs2.2", "resblocks.9.convs2.0", "resblocks.9.convs2.1", "resblocks.9.convs2.2"):
                print("   ", r["name"], r["kernel"], round(r["ms"], 4))
