#!/usr/bin/env python
"""One warm-up forward + one profiled forward of the bench workload (HiFi-GAN V1, 16 x 800 frames),
for ncu.  No torch kernels are launched: every launch ncu sees is one of ours — 61 per forward in bf16
mode (mel_to_operand, the tcgen05 kernels conv_tc2_kernel / conv_pair_tc_kernel / conv_tc_kernel,
conv_post32_kernel), 79 in fp32 mode (no pair fusion, no CTA-pair kernel).

    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 61 --launch-count 61 \
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
