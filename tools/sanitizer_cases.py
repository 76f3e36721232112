#!/usr/bin/env python
"""Small cases that touch every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitizer_cases.py
    compute-sanitizer --tool synccheck python tools/sanitizer_cases.py

dense + ragged forwards of V1 / V2-style / V3-style (tcgen05 single-CTA, CTA-pair, fused pair, narrow,
CUDA-core, conv_post), both precisions, and the PostNet / mel_linear stack."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from tts_king_b200.fs_two.model.fastspeech2 import MelLinear, mel_tail  # noqa: E402
from tts_king_b200.fs_two.transformer.Layers import PostNet  # noqa: E402
from _util import make_generator  # noqa: E402

with torch.no_grad():
    for name, cfg in (("v1", fx.V1), ("v2_narrow", fx.V2_NARROW), ("v3_rb2", fx.V3_RB2)):
        for prec in ("bf16", "fp32"):
            m = make_generator(cfg, precision=prec).cuda()
            mel = fx.synthetic_mel(3, 37, seed=7).cuda()
            y = m(mel)
            yr = m.forward_ragged(mel, [37, 5, 20])
            yi = m.generate_int16(mel, frames=[1, 37, 36])
            torch.cuda.synchronize()
            ok = bool(torch.isfinite(y).all()) and bool(torch.equal(yr[1, 0, :5 * m.hop_length], y[1, 0, :5 * m.hop_length]))
            print(name, prec, "ok" if ok else "MISMATCH", tuple(y.shape), tuple(yi.shape))
    # the throughput schedule (enough tiles for the CTA-pair kernels and the 256-column N tiles); with HG_FOLD=2 in the
    # environment every fused pair runs on the time-folded kernel, with HG_FOLD=0 on the N = C kernel
    m = make_generator(fx.V1, precision="bf16").cuda()
    mel = fx.synthetic_mel(2, 300, seed=9).cuda()
    y = m(mel)
    yr = m.forward_ragged(mel, [300, 41])
    torch.cuda.synchronize()
    print("v1 bf16 2x300", "ok" if bool(torch.isfinite(y).all()) and bool(torch.equal(yr[1, 0, :41 * 256], y[1, 0, :41 * 256])) else "MISMATCH")
    for prec in ("bf16", "fp32", "fp32_ffma"):
        post = PostNet(**fx.POSTNET_TINY, precision=prec).eval().cuda()
        lin = MelLinear(48, 80, precision=prec).cuda()
        out, pout = mel_tail(torch.randn(2, 29, 48).cuda(), lin, post)
        torch.cuda.synchronize()
        print("postnet", prec, "ok" if bool(torch.isfinite(pout).all()) else "NONFINITE", tuple(pout.shape))
