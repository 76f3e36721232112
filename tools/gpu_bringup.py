#!/usr/bin/env python
"""GPU bring-up driver: runs independent check groups in subprocesses (a trapped kernel poisons its
CUDA context, not the others) and writes gpurun_out/bringup_*.log.

    python tools/gpu_bringup.py                 # all groups
    python tools/gpu_bringup.py ops_bf16 e2e    # selected groups
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")

GROUPS = ["selftest", "ops_fp32_ffma", "ops_bf16", "ops_fp32", "e2e_fp32_ffma", "e2e_bf16", "e2e_fp32", "perf"]


def g_selftest():
    from tts_king_b200 import _native
    n, rep = _native.selftest(0)
    print(rep)
    print("mismatching cases:", n)


def _ops(prec):
    import torch
    import test_gpu_ops as T
    torch.manual_seed(0)
    for (C, k, d, n) in [(64, 3, 1, 300), (64, 7, 3, 300), (32, 3, 1, 300), (32, 11, 5, 301), (128, 7, 3, 301),
                         (128, 11, 5, 700), (256, 3, 1, 301), (256, 11, 5, 301), (64, 7, 3, 1), (64, 7, 3, 129)]:
        g = torch.Generator().manual_seed(C + k + d)
        x = torch.randn(2, C, n, generator=g)
        w = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
        b = torch.randn(C, generator=g) * 0.1
        res = torch.randn(2, C, n, generator=g)
        try:
            y = T.run_conv1d(x, w, b, d, 0.1, res, prec)
            ref = T.ref_conv1d(x, w, b, d, 0.1, res, prec)
            err = (y.double() - ref).abs()
            rel = err.max().item() / ref.abs().max().item()
            bad = (err > 1e-3 * ref.abs().max()).float().mean().item()
            print(f"conv1d {prec} C={C} k={k} d={d} L={n}: rel_max_err={rel:.3e} frac_bad={bad:.4f} nan={int(torch.isnan(y).sum())}",
                  "OK" if rel < 1e-4 else "FAIL")
            if rel >= 1e-4:
                e2 = err[0].max(dim=0).values  # per time step
                idx = (e2 > 1e-3 * ref.abs().max()).nonzero().flatten()
                print("   bad time steps (item 0):", idx[:20].tolist(), "... count", idx.numel())
                e3 = err[0].max(dim=1).values
                idc = (e3 > 1e-3 * ref.abs().max()).nonzero().flatten()
                print("   bad channels (item 0):", idc[:20].tolist(), "... count", idc.numel())
        except Exception as ex:  # noqa
            print(f"conv1d {prec} C={C} k={k} d={d} L={n}: EXCEPTION {ex}")
    import torch.nn.functional as F
    from tts_king_b200 import _native
    L = _native.lib()
    for (cin, cout, k, s) in [(64, 32, 4, 2), (128, 64, 4, 2), (256, 128, 16, 8), (512, 256, 16, 8)]:
        g = torch.Generator().manual_seed(cin)
        x = torch.randn(2, cin, 37, generator=g)
        w = torch.randn(cin, cout, k, generator=g) / (cin * k / s) ** 0.5
        b = torch.randn(cout, generator=g) * 0.1
        try:
            xc = x.transpose(1, 2).contiguous().cuda()
            y = torch.full((2, 37 * s, cout), float("nan"), device="cuda")
            _native.check(L.hg_op_conv_transpose1d(0, T.PREC[prec], xc.data_ptr(), 2, 37, cin, w.data_ptr(), b.data_ptr(),
                                                   cout, k, s, 0.1, y.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            y = y.cpu().transpose(1, 2)
            ref = F.conv_transpose1d(T.operand_model(F.leaky_relu(x, 0.1), prec), T.operand_model(w, prec), b.double(),
                                     stride=s, padding=(k - s) // 2)
            rel = (y.double() - ref).abs().max().item() / ref.abs().max().item()
            print(f"convT {prec} {cin}->{cout} k={k} s={s}: rel_max_err={rel:.3e} nan={int(torch.isnan(y).sum())}",
                  "OK" if rel < 1e-4 else "FAIL")
        except Exception as ex:  # noqa
            print(f"convT {prec} {cin}->{cout}: EXCEPTION {ex}")


def _e2e(prec):
    import numpy as np
    import torch
    from oracle import fixtures as fx
    from oracle.common import max_abs, snr_db, ac_snr_db
    from _util import golden, make_generator, stored_state
    for name, cfg in (("tiny_rb1", fx.TINY_RB1), ("tiny_rb2", fx.TINY_RB2)):
        g = golden(name)
        m = make_generator(cfg, seed=5, fold=False, precision=prec)
        m.load_state_dict(stored_state(g))
        m.cuda()
        with torch.no_grad():
            y = m(torch.from_numpy(g["mel"]).cuda()).cpu().numpy()
        print(f"e2e {prec} {name}: max_abs={max_abs(y, g['y']):.3e} snr={snr_db(g['y'], y):.1f} dB")
    for name, cfg in (("v1", fx.V1), ("v2_narrow", fx.V2_NARROW), ("v3_rb2", fx.V3_RB2)):
        g = golden(name + "_seed1234")
        m = make_generator(cfg, precision=prec).cuda()
        with torch.no_grad():
            y = m(torch.from_numpy(g["mel_a"]).cuda()).cpu().numpy()
            yb = m(torch.from_numpy(g["mel_b"]).cuda()).cpu().numpy()
        print(f"e2e {prec} {name}: max_abs={max_abs(y, g['y_a']):.3e} snr={snr_db(g['y_a'], y):.1f} dB "
              f"ac_snr={ac_snr_db(g['y_a'], y):.1f} dB | mel_b max_abs={max_abs(yb, g['y_b']):.3e}")
        m.load_state_dict(fx.alive_state(cfg))
        with torch.no_grad():
            ya = m(torch.from_numpy(g["mel_a"]).cuda()).cpu().numpy()
        print(f"e2e {prec} {name} alive: max_abs={max_abs(ya, g['y_alive_a']):.3e} snr={snr_db(g['y_alive_a'], ya):.1f} dB")


def g_perf():
    import torch
    from oracle import fixtures as fx
    from _util import make_generator
    m = make_generator(fx.V1, precision="bf16").cuda()
    mel = fx.synthetic_mel(16, 800, seed=7).cuda()
    for prec in ("bf16", "fp32"):
        m.precision = prec
        with torch.no_grad():
            for _ in range(2):
                m(mel)
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(3):
                m(mel)
            t1.record(); torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 3
        audio_s = 16 * 800 * 256 / 22050
        print(f"perf {prec}: {ms:.2f} ms/step  {audio_s / ms * 1e3:.0f} audio-s/s  {7.86e12 / ms / 1e9:.1f} TFLOP/s")
        rows = m.profile_layers(mel)
        tot = sum(r["ms"] for r in rows)
        print(f"  profiled total {tot:.2f} ms over {len(rows)} launches")
        for r in rows:
            if r["kind"] < 0:
                print(f"  {r['name']:28s} {r['ms']:.3f} ms")
                continue
            print(f"  {r['name']:28s} {r['c_in']:4d}->{r['c_out']:4d} k={r['k']:2d} d={r['dilation']} s={r['stride']} "
                  f"tc={int(r['tensor_core'])} nt={r['n_tile']} ms={r['m_subtiles']} st={r['stages']} res={int(r['weights_resident'])} nb={r['slab_buffers']} smem={r['smem_bytes']} "
                  f"{r['ms']:.3f} ms")
        json.dump(rows, open(os.path.join(OUT, f"layers_{prec}.json"), "w"))


def run_group(name):
    if name == "selftest":
        g_selftest()
    elif name.startswith("ops_"):
        _ops(name[4:])
    elif name.startswith("e2e_"):
        _e2e(name[4:])
    elif name == "perf":
        g_perf()
    else:
        raise SystemExit("unknown group " + name)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        run_group(sys.argv[2])
        sys.exit(0)
    groups = sys.argv[1:] or GROUPS
    for gname in groups:
        t = time.time()
        log = os.path.join(OUT, f"bringup_{gname}.log")
        with open(log, "w") as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", gname], stdout=f,
                                   stderr=subprocess.STDOUT, timeout=600)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "timeout"
        print(f"== {gname}: rc={rc} {time.time() - t:.1f}s")
        print(open(log).read()[-6000:])
