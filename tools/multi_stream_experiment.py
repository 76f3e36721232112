#!/usr/bin/env python
"""Does running independent sub-batches on several CUDA streams fill the tails of the per-layer
kernels?  (Batch items never interact; each layer is one persistent kernel whose last round of tiles
leaves SMs idle, and consecutive layers of one stream cannot overlap because of their halo dependency.)

    python tools/multi_stream_experiment.py [B] [T]
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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
m = make_generator(fx.V1, precision="bf16").cuda()
mel = fx.synthetic_mel(B, T, seed=7).cuda()
res = {"B": B, "T": T}
with torch.no_grad():
    ref = m(mel)
    for n_streams in (1, 2, 4):
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        parts = torch.chunk(mel, n_streams, dim=0)

        def step():
            outs = []
            cur = torch.cuda.current_stream()
            for s, x in zip(streams, parts):
                s.wait_stream(cur)
                with torch.cuda.stream(s):
                    outs.append(m(x))
            for s in streams:
                cur.wait_stream(s)
            return outs

        for _ in range(3):
            outs = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        e0.record()
        for _ in range(reps):
            outs = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        same = bool(torch.equal(torch.cat(outs, 0), ref))
        res[f"streams_{n_streams}"] = {"ms": ms, "audio_s_per_s": B * T * 256 / 22050 / ms * 1e3, "bitwise_equal_to_single": same}
print(json.dumps(res))
