#!/usr/bin/env python
"""EXPERIMENT (timing only, results are wrong by construction): per-launch time of the time-folded pair kernels
with parts of their work switched off through HG_FOLD_DBG (conv_pair_fold.cu, TcFoldParams::dbg):

    1 no MMAs   2 no global stores   4 no residual TMA   8 no slab TMA   16 no staging read-modify-write   32 no xt stores
    64 no weight streaming (ring kernels)   256 no weight waits / producer   512 no weight-stage releases
    1024 epilogue bodies off (hand-shakes only)

The switches are compiled in only with -DHG_FOLD_DBG:

    HG_NVCC_EXTRA=-DHG_FOLD_DBG python -m tts_king_b200.build --force && python tools/fold_dbg.py

(rebuild without it afterwards).  Results: profiles/r2_pair_epilogue_issue_bound.md.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

MODES = [int(x) for x in os.environ.get("MODES", "0,1,2,4,8,16,32,3,6,12,14,15,63,62").split(",")]
ROUNDS = int(os.environ.get("ROUNDS", "4"))


def main():
    os.environ["HG_FOLD"] = os.environ.get("HG_FOLD", "2")
    mel = fx.synthetic_mel(16, 800, seed=7).cuda()
    m = make_generator(fx.V1, precision="bf16").cuda()
    res = {}
    with torch.no_grad():
        for _ in range(3):
            m(mel)
        for _ in range(ROUNDS):
            for mode in MODES:
                os.environ["HG_FOLD_DBG"] = str(mode)
                for r in m.profile_layers(mel):
                    if r.get("kernel") == "tcgen05 fused pair":
                        res.setdefault(r["name"], {}).setdefault(mode, []).append(r["ms"])
        os.environ["HG_FOLD_DBG"] = "0"
    print("launch | " + " | ".join(str(x) for x in MODES))
    for ln in sorted(res):
        print(ln, "|", " | ".join(f"{sum(res[ln][x]) / len(res[ln][x]):.3f}" for x in MODES))


if __name__ == "__main__":
    main()
