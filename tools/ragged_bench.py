#!/usr/bin/env python
"""N2 measurement: a padded synth batch (B utterances, lengths uniform in [T/4, T]) through
vocoder_infer with and without `lengths` (reference call site fs_two/utils/tools.py:257-268).

    python tools/ragged_bench.py [B] [T] > gpurun_out/ragged_bench.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import fixtures as fx  # noqa: E402  (synthetic inputs only)
from tts_king_b200 import parallel, ragged  # noqa: E402
from tts_king_b200.fs_two.utils.model import vocoder_infer  # noqa: E402
from _util import make_generator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 800
m = make_generator(fx.V1, precision="bf16").cuda()
rng = np.random.default_rng(0)
frames = [int(v) for v in rng.integers(T // 4, T + 1, size=B)]
frames[0] = T
mel = fx.synthetic_mel(B, T, seed=3).pin_memory()
lengths = torch.tensor([f * 256 for f in frames])
mc = {"vocoder": {"model": "HiFi-GAN"}}
pc = {"preprocessing": {"audio": {"max_wav_value": 32768.0}}}


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


res = {"B": B, "T": T, "frames": frames, "valid_fraction": sum(frames) / (B * T)}
res["padded_ms"] = timed(lambda: vocoder_infer(mel, m, mc, pc))
halo = parallel.halo_frames(m.h)
for lc in (100, 200, 400, 800, 1600):
    buckets = ragged.plan_length_buckets(frames, halo, T, lc)
    md = mel.cuda()
    t = timed(lambda: ragged.ragged_generate(m, md, lengths.tolist(), launch_cost=lc, mode="buckets"))
    res[f"buckets_device_ms_launch_cost_{lc}"] = {"ms": t, "buckets": [[len(b), ragged.bucket_extent(frames, b, halo, T)] for b in buckets]}
md = mel.cuda()
res["padded_device_ms"] = timed(lambda: m.generate_int16(md))
res["kernel_ragged_device_ms"] = timed(lambda: m.generate_int16(md, frames=frames))
res["ideal_device_ms"] = res["padded_device_ms"] * sum(min(T, f + halo) for f in frames) / (B * T)
res["ragged_ms"] = timed(lambda: vocoder_infer(mel, m, mc, pc, lengths=lengths))
res["speedup_e2e"] = res["padded_ms"] / res["ragged_ms"]
a = vocoder_infer(mel, m, mc, pc)
b = vocoder_infer(mel, m, mc, pc, lengths=lengths)
res["bit_identical"] = all(np.array_equal(x[:len(y)], y) for x, y in zip(a, b))
print(json.dumps(res))
