#!/usr/bin/env python
"""One tiny V1 forward per precision (every tcgen05 kernel family + conv_post) for
`compute-sanitizer --tool racecheck` (shared-memory hazards)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

with torch.no_grad():
    for prec in ("bf16", "fp32"):
        m = make_generator(fx.V1, precision=prec).cuda()
        y = m(fx.synthetic_mel(2, 9, seed=7).cuda())
        torch.cuda.synchronize()
        print(prec, "ok" if bool(torch.isfinite(y).all()) else "NONFINITE", tuple(y.shape))
