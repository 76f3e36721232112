#!/usr/bin/env python
"""A/B of the fused ResBlock-pair kernels on the bench workload (16 x 800 frames, bf16), interleaved in ONE
process so that clock / power drift hits every variant alike: three generators are built under different
HG_FOLD settings (read at plan creation) and timed round-robin.

    N=C      HG_FOLD=0   conv_pair_tc.cu everywhere
    fold     HG_FOLD=2   conv_pair_fold.cu wherever it applies
    default              the shipped mix (api.cu::fold_pays)

Prints whole-forward time per variant (mean of the rounds), per-launch times of the 18 pair launches (mean of
the per-launch event passes) and the sha256 of the waveform (all variants must give the same bits).
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from oracle import fixtures as fx  # noqa: E402
from _util import make_generator  # noqa: E402

VARIANTS = [("N=C", {"HG_FOLD": "0"}), ("fold", {"HG_FOLD": "2"}), ("default", {})]
if os.environ.get("VARIANTS"):  # e.g. VARIANTS="forward:HG_TILE_ORDER=0;alternate:" — any plan-creation switches
    VARIANTS = []
    for item in os.environ["VARIANTS"].split(";"):
        name, _, kv = item.partition(":")
        VARIANTS.append((name, dict(x.split("=") for x in kv.split(",") if x)))
ROUNDS = int(os.environ.get("ROUNDS", "6"))


def main():
    mel = fx.synthetic_mel(16, 800, seed=7).cuda()
    models, digests = {}, {}
    for name, env in VARIANTS:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        m = make_generator(fx.V1, precision=os.environ.get("PRECISION", "bf16")).cuda()
        with torch.no_grad():
            y = m(mel)  # the plan (and its HG_* switches) is created here
        torch.cuda.synchronize()
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        models[name] = m
        digests[name] = hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]
    step = {n: [] for n in models}
    pairs = {n: {} for n in models}
    with torch.no_grad():
        for name, m in models.items():
            for _ in range(3):
                m(mel)
        names = list(models)
        for rnd in range(ROUNDS):
            for name in names[rnd % len(names):] + names[:rnd % len(names)]:  # rotate: nobody always follows the same variant
                m = models[name]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    m(mel)
                e1.record()
                torch.cuda.synchronize()
                step[name].append(e0.elapsed_time(e1) / 10)
                for r in m.profile_layers(mel):
                    if r.get("kernel") == "tcgen05 fused pair":
                        pairs[name].setdefault(r["name"], []).append(r["ms"])
    mean = lambda v: sum(v) / len(v)  # noqa: E731
    print("variant | forward ms (mean of %d x 10) | min | 18 pair launches, sum of means ms | digest" % ROUNDS)
    for n in models:
        print(f"{n} | {mean(step[n]):.3f} | {min(step[n]):.3f} | {sum(mean(v) for v in pairs[n].values()):.3f} | {digests[n]}")
    print("launch | " + " | ".join(models))
    for ln in sorted(pairs[VARIANTS[0][0]]):
        print(ln, "|", " | ".join(f"{mean(pairs[n][ln]):.4f}" for n in models))
    assert len(set(digests.values())) == 1, digests


if __name__ == "__main__":
    main()
