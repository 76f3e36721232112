#!/usr/bin/env python
"""A/B of the fused ResBlock-pair kernels on the bench workload (16 x 800 frames, bf16): per-launch times of
the 18 pair launches for the N = C kernel (HG_FOLD=0) and the time-folded kernel with each E2 variant
(HG_FOLD_E2=0/1/2), plus a sha256 of the waveform (all variants must give the same bits).

    python tools/pair_modes.py                 # all variants, one subprocess each
    HG_TC_DEBUG_TIMING=resblocks.8.convs2.0 python tools/pair_modes.py --one   # wait breakdown of one launch
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one():
    import torch

    from oracle import fixtures as fx
    from _util import make_generator

    m = make_generator(fx.V1, precision="bf16").cuda()
    mel = fx.synthetic_mel(16, 800, seed=7).cuda()
    with torch.no_grad():
        y = m(mel)
        torch.cuda.synchronize()
        digest = hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16]
        for _ in range(3):
            m(mel)
        rows = m.profile_layers(mel)
        rows = m.profile_layers(mel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m(mel)
        e1.record()
        torch.cuda.synchronize()
    pairs = {r["name"]: round(r["ms"], 4) for r in rows if r.get("kernel") == "tcgen05 fused pair"}
    print(json.dumps({"digest": digest, "ms_per_step": e0.elapsed_time(e1) / 20, "pair_ms_total": round(sum(pairs.values()), 3),
                      "pairs": pairs}))


def main():
    if "--one" in sys.argv:
        return one()
    variants = [("N=C kernel (HG_FOLD=0)", {"HG_FOLD": "0"}), ("fold everywhere (HG_FOLD=2)", {"HG_FOLD": "2"}), ("default mix", {})]
    out = {}
    for name, env in variants:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], capture_output=True, text=True,
                           env={**os.environ, **env}, timeout=600)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print(name, "FAILED", r.stderr[-1500:])
            continue
        out[name] = json.loads(line[-1])
        if r.stderr.strip():
            print(r.stderr.strip()[-3000:])
    names = sorted(next(iter(out.values()))["pairs"]) if out else []
    print("variant | step ms | pairs total ms | digest")
    for name, d in out.items():
        print(f"{name} | {d['ms_per_step']:.3f} | {d['pair_ms_total']} | {d['digest']}")
    print("launch | " + " | ".join(out))
    for n in names:
        print(n, "|", " | ".join(str(d["pairs"].get(n)) for d in out.values()))


if __name__ == "__main__":
    main()
