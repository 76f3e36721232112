#!/usr/bin/env python
"""N3 measurement: mel_linear + PostNet (+ residual) on the GPU for the bench batch (16 x 800 frames),
per precision, next to the CPU oracle on a bounded sample, and the whole tail -> vocoder chain.

    python tools/postnet_bench.py > gpurun_out/postnet_bench.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import fixtures as fx  # noqa: E402
from oracle import postnet_oracle as po  # noqa: E402  (CPU baseline leg only)
from tts_king_b200.fs_two.model.fastspeech2 import MelLinear, mel_tail, mel_to_vocoder  # noqa: E402
from tts_king_b200.fs_two.transformer.Layers import PostNet  # noqa: E402
from _util import make_generator  # noqa: E402

B, T = 16, 800
FLOP_PER_FRAME = 2 * (256 * 80 + 5 * (80 * 512 + 3 * 512 * 512 + 512 * 80))


def gpu_ms(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {"B": B, "T": T, "gflop": FLOP_PER_FRAME * B * T / 1e9}
torch.manual_seed(1234)
sd = fx.alive_batchnorm_({k: v.clone() for k, v in PostNet(**fx.POSTNET_FULL).state_dict().items()})
dec = torch.randn(B, T, 256, generator=torch.Generator().manual_seed(1)).cuda()
for prec in ("fp32", "bf16"):
    post = PostNet(**fx.POSTNET_FULL, precision=prec)
    post.load_state_dict(sd)
    post.eval().cuda()
    torch.manual_seed(77)
    lin = MelLinear(256, 80, precision=prec).cuda()
    ms = gpu_ms(lambda: mel_tail(dec, lin, post))
    res[prec] = {"ms": ms, "tflops": FLOP_PER_FRAME * B * T / ms / 1e9, "frames_per_s": B * T / ms * 1e3}
    if prec == "bf16":
        gen = make_generator(fx.V1, precision="bf16").cuda()
        res["tail_plus_vocoder_bf16_ms"] = gpu_ms(lambda: gen.generate_int16(mel_to_vocoder(mel_tail(dec, lin, post)[1])), n=10)
        md = mel_tail(dec, lin, post)[1].transpose(1, 2).contiguous()
        res["vocoder_only_bf16_ms"] = gpu_ms(lambda: gen.generate_int16(md), n=10)
# CPU oracle (the reference's own ATen calls), bounded sample: 1 x 800 frames
lw, lb = torch.randn(80, 256) * 0.05, torch.zeros(80)
xs = dec[:1].cpu()
torch.set_num_threads(os.cpu_count() or 1)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    with torch.no_grad():
        po.mel_tail(lw, lb, sd, xs)
    best = min(best, time.perf_counter() - t0)
res["cpu_oracle"] = {"frames_per_s": T / best, "cores": torch.get_num_threads(), "sample": "1 x 800 frames, best of 5"}
print(json.dumps(res))
