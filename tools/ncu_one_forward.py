#!/usr/bin/env python
"""One warm-up forward + one profiled forward of the bench workload (HiFi-GAN V1, 16 x 800 frames),
for ncu.  No torch kernels are launched: every launch ncu sees is one of ours (79 per forward:
mel_to_operand, 77 x conv_tc_kernel, conv_post_kernel).

    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 79 --launch-count 79 \
        --csv --log-file gpurun_out/launches.csv python tools/ncu_one_forward.py [bf16|fp32]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402  (input/weight recipe only)
from _util import make_generator  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(os.environ.get("HG_B", "16"))
T = int(os.environ.get("HG_T", "800"))
m = make_generator(fx.V1, precision=prec).cuda()
mel = fx.synthetic_mel(B, T, seed=7).cuda()
with torch.no_grad():
    m(mel)
    torch.cuda.synchronize()
    y = m(mel)
    torch.cuda.synchronize()
print("ok", tuple(y.shape))
