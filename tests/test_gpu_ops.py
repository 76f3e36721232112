"""T1/T2 — op- and block-level parity of the CUDA kernels (through the C ABI) against fp64
torch.nn.functional on the host.  Needs a B200: run with `-m gpu` under gpurun."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tts_king_b200 import _native

pytestmark = pytest.mark.gpu

PREC = {"fp32": 1, "bf16": 0, "fp32_ffma": 2}


def bf16_round(t):
    return t.float().bfloat16().double()


def split_round(t):
    hi = t.float().bfloat16().float()
    lo = (t.float() - hi).bfloat16().float()
    return (hi + lo).double()


def operand_model(t, prec):
    """What the kernel contracts over, given fp32 values: exact, bf16, or hi+lo."""
    return {"fp32_ffma": t.double(), "bf16": bf16_round(t), "fp32": split_round(t)}[prec]


def run_conv1d(x, w, b, d, slope, res, prec):
    """x [B,C,L] cpu float -> y [B,Cout,L] via hg_op_conv1d (channels-last on device)."""
    L = _native.lib()
    dev = torch.device("cuda", 0)
    xc = x.transpose(1, 2).contiguous().to(dev)
    rc = res.transpose(1, 2).contiguous().to(dev) if res is not None else None
    B, n, cin = xc.shape
    cout, _, k = w.shape
    y = torch.full((B, n, cout), float("nan"), device=dev)
    wc, bc = w.contiguous(), b.contiguous()
    _native.check(L.hg_op_conv1d(0, PREC[prec], xc.data_ptr(), B, n, cin, wc.data_ptr(), bc.data_ptr(), cout, k, d,
                                 float(slope), rc.data_ptr() if rc is not None else None, y.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return y.cpu().transpose(1, 2)


def ref_conv1d(x, w, b, d, slope, res, prec):
    a = operand_model(F.leaky_relu(x, slope) if slope != 1.0 else x, prec)
    ww = operand_model(w, prec) if prec != "fp32" else split_round(w)
    k = w.shape[-1]
    y = F.conv1d(a, ww, b.double(), dilation=d, padding=(k * d - d) // 2)
    return y + res.double() if res is not None else y


TOL = {"fp32_ffma": 2e-6, "fp32": 3e-5, "bf16": 3e-5}  # relative to max|y|, vs the operand model


@pytest.mark.parametrize("prec", ["fp32_ffma", "fp32", "bf16"])
@pytest.mark.parametrize("C,k,d", [(256, 3, 1), (256, 11, 5), (128, 7, 3), (128, 11, 1), (64, 3, 5), (64, 11, 5),
                                   (32, 3, 1), (32, 7, 3), (32, 11, 5)])
def test_conv1d_parity(C, k, d, prec):
    g = torch.Generator().manual_seed(C * 100 + k * 10 + d)
    B, n = 2, 301  # odd length: ragged last tile, and shorter than one 2x128-row CTA tile pair
    x = torch.randn(B, C, n, generator=g)
    w = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b = torch.randn(C, generator=g) * 0.1
    res = torch.randn(B, C, n, generator=g)
    y = run_conv1d(x, w, b, d, 0.1, res, prec)
    ref = ref_conv1d(x, w, b, d, 0.1, res, prec)
    err = (y.double() - ref).abs().max().item()
    assert not torch.isnan(y).any()
    assert err <= TOL[prec] * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("n", [1, 2, 127, 128, 129, 255, 256, 257, 1000])
def test_conv1d_lengths(n, prec):
    """Tile-boundary lengths and T = 1."""
    g = torch.Generator().manual_seed(n)
    x = torch.randn(3, 64, n, generator=g)
    w = torch.randn(64, 64, 7, generator=g) / 21.0
    b = torch.randn(64, generator=g)
    y = run_conv1d(x, w, b, 3, 0.1, None, prec)
    ref = ref_conv1d(x, w, b, 3, 0.1, None, prec)
    assert (y.double() - ref).abs().max().item() <= TOL[prec] * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("prec", ["fp32_ffma", "fp32", "bf16"])
@pytest.mark.parametrize("C,k,d,n", [(16, 3, 1, 700), (16, 11, 5, 513), (16, 7, 3, 1), (8, 3, 5, 1025), (8, 11, 5, 300),
                                     (8, 7, 1, 511)])
def test_narrow_conv1d_parity(C, k, d, n, prec):
    """The CUDA-core kernel of the 16- / 8-channel stages (conv_narrow.cu, BASELINE cfg-4): exact fp32
    arithmetic on whatever operand format the mode stores."""
    g = torch.Generator().manual_seed(C * 100 + k * 10 + d)
    x = torch.randn(2, C, n, generator=g)
    w = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b = torch.randn(C, generator=g) * 0.1
    res = torch.randn(2, C, n, generator=g)
    y = run_conv1d(x, w, b, d, 0.1, res, prec)
    a = operand_model(F.leaky_relu(x, 0.1), prec)  # activations are stored in the mode's format; weights stay fp32
    ref = F.conv1d(a, w.double(), b.double(), dilation=d, padding=(k * d - d) // 2) + res.double()
    assert not torch.isnan(y).any()
    assert (y.double() - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


def test_conv1d_fp32_mode_is_fp32_accurate():
    """HG_PREC_FP32 (bf16x3) against the true fp64 result of the fp32 inputs — the accuracy the
    north_star's 1e-4 end-to-end bound rests on."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 128, 500, generator=g)
    w = torch.randn(128, 128, 11, generator=g) / (128 * 11) ** 0.5
    b = torch.zeros(128)
    y = run_conv1d(x, w, b, 5, 1.0, None, "fp32")
    ref = F.conv1d(x.double(), w.double(), None, dilation=5, padding=25)
    rel = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert rel <= 5e-5, rel


@pytest.mark.parametrize("prec", ["fp32_ffma", "fp32", "bf16"])
@pytest.mark.parametrize("cin,cout,k,s", [(512, 256, 16, 8), (256, 128, 16, 8), (128, 64, 4, 2), (64, 32, 4, 2),
                                          (256, 128, 8, 4), (16, 16, 4, 2), (16, 16, 16, 8), (16, 8, 4, 2)])
def test_conv_transpose1d_parity(cin, cout, k, s, prec):
    L = _native.lib()
    g = torch.Generator().manual_seed(cin + k)
    B, n = 2, 37
    x = torch.randn(B, cin, n, generator=g)
    w = torch.randn(cin, cout, k, generator=g) / (cin * k / s) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    dev = torch.device("cuda", 0)
    xc = x.transpose(1, 2).contiguous().to(dev)
    y = torch.full((B, n * s, cout), float("nan"), device=dev)
    _native.check(L.hg_op_conv_transpose1d(0, PREC[prec], xc.data_ptr(), B, n, cin, w.data_ptr(), b.data_ptr(), cout, k,
                                           s, 0.1, y.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    y = y.cpu().transpose(1, 2)
    a = operand_model(F.leaky_relu(x, 0.1), prec)
    # 16 input channels have no tensor-core tiling: CUDA-core kernels, whose weights stay fp32 (convt_narrow16 in
    # bf16 mode when C_out = 16 too — the upsampler in front of a stage padded to 16 channels — conv_ffma otherwise)
    ww = operand_model(w, prec) if cin != 16 else w.double()
    ref = F.conv_transpose1d(a, ww, b.double(), stride=s, padding=(k - s) // 2)
    assert not torch.isnan(y).any()
    assert (y.double() - ref).abs().max().item() <= TOL[prec] * ref.abs().max().item()


@pytest.mark.parametrize("C", [32, 64, 8, 2])
def test_conv_post_parity(C):
    L = _native.lib()
    g = torch.Generator().manual_seed(C)
    B, n = 3, 777  # 777 = 3 tiles of 256 + 9: ragged last tile
    x = torch.randn(B, C, n, generator=g)
    w = torch.randn(1, C, 7, generator=g) / (7 * C) ** 0.5
    b = torch.randn(1, generator=g)
    dev = torch.device("cuda", 0)
    xc = x.transpose(1, 2).contiguous().to(dev)
    y = torch.full((B, n), float("nan"), device=dev)
    _native.check(L.hg_op_conv_post(0, xc.data_ptr(), B, n, C, w.data_ptr(), b.data_ptr(), y.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = torch.tanh(F.conv1d(F.leaky_relu(x.double(), 0.01), w.double(), b.double(), padding=3))[:, 0]
    assert (y.cpu().double() - ref).abs().max().item() <= 2e-6


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_resblock1_block_level(prec):
    """T2: ResBlock1(C,k,(1,3,5)) stand-alone against the fp64 restatement of hifi/models.py:88-95."""
    from oracle import fixtures as fx
    from tts_king_b200.hifi.models import ResBlock1

    torch.manual_seed(3)
    blk = ResBlock1(fx.make_h(fx.V1), 128, 7, (1, 3, 5))
    blk.precision = prec
    x = torch.randn(2, 128, 333)
    y = blk.cuda()(x.cuda()).cpu()
    r = x.double()
    for c1, c2 in zip(blk.convs1, blk.convs2):
        w1 = torch._weight_norm(c1.weight_v, c1.weight_g, 0).detach().cpu().double()
        w2 = torch._weight_norm(c2.weight_v, c2.weight_g, 0).detach().cpu().double()
        xt = F.conv1d(F.leaky_relu(r, 0.1), w1, c1.bias.detach().cpu().double(), dilation=c1.dilation[0], padding=c1.padding[0])
        xt = F.conv1d(F.leaky_relu(xt, 0.1), w2, c2.bias.detach().cpu().double(), padding=c2.padding[0])
        r = xt + r
    err = (y.double() - r).abs().max().item()
    assert err <= (2e-5 if prec == "fp32" else 5e-2) * r.abs().max().item(), err


@pytest.mark.parametrize("C,k,d1", [(64, 3, 1), (64, 7, 3), (64, 11, 5), (32, 3, 5), (32, 7, 1), (32, 11, 5), (64, 11, 1), (32, 11, 3),
                                    (16, 3, 1), (16, 7, 3), (16, 11, 5), (16, 3, 5), (16, 11, 1)])
@pytest.mark.parametrize("n", [1, 100, 246, 247, 502, 503, 1500, 8, 240, 244, 248, 460, 480, 496, 920, 1000, 1996, 4104])
def test_fused_pair_parity(C, k, d1, n):
    """The fused ResBlock-pair kernels against the fp64 restatement of hifi/models.py:90-94 with the
    kernels' operand model (bf16 operands, bf16 intermediate).  Lengths that are a multiple of F = 128 / C
    run the time-folded kernel (conv_pair_fold.cu; output tiles of 230..246 rows at C = 64, 460..496 at
    C = 32, 920..1000 at C = 16, and a zeroed tail inside the last F*d1-row block group unless F*d1 divides
    the length), the others the N = C kernel (conv_pair_tc.cu, 246/502-row tiles); both are straddled.
    Where both kernels apply they must agree bit for bit (same summation order for every output)."""
    if C == 16 and n % 8:
        pytest.skip("16 channels: only the time-folded kernel exists, and it needs L % 8 == 0")
    L = _native.lib()
    g = torch.Generator().manual_seed(C * 1000 + k * 10 + d1 + n)
    B = 2
    x = torch.randn(B, C, n, generator=g)
    w1 = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b1 = torch.randn(C, generator=g) * 0.1
    w2 = torch.randn(C, C, k, generator=g) / (C * k) ** 0.5
    b2 = torch.randn(C, generator=g) * 0.1
    res = torch.randn(B, C, n, generator=g)
    dev = torch.device("cuda", 0)
    xc = x.transpose(1, 2).contiguous().to(dev)
    rc_ = res.transpose(1, 2).contiguous().to(dev)
    y = torch.full((B, n, C), float("nan"), device=dev)
    _native.check(L.hg_op_conv_pair(0, xc.data_ptr(), B, n, C, k, d1, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                    b2.data_ptr(), 0.1, rc_.data_ptr(), y.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if C != 16:
        import os
        y_nc = torch.full((B, n, C), float("nan"), device=dev)
        os.environ["HG_FOLD"] = "0"
        try:
            _native.check(L.hg_op_conv_pair(0, xc.data_ptr(), B, n, C, k, d1, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                            b2.data_ptr(), 0.1, rc_.data_ptr(), y_nc.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
        finally:
            del os.environ["HG_FOLD"]
        torch.cuda.synchronize()
        assert torch.equal(y, y_nc)
    y = y.cpu().transpose(1, 2)
    a = bf16_round(F.leaky_relu(x, 0.1))
    xt = F.conv1d(a, bf16_round(w1), b1.double(), dilation=d1, padding=(k * d1 - d1) // 2)
    xt = bf16_round(F.leaky_relu(xt.float(), 0.1))  # the kernel rounds fp32 accum + bias, then lrelu, to bf16
    ref = F.conv1d(xt, bf16_round(w2), b2.double(), padding=(k - 1) // 2) + res.double()
    assert not torch.isnan(y).any()
    # the intermediate's bf16 rounding can flip on fp32-vs-fp64 accumulation noise: allow a few bf16 ulps of xt
    assert (y.double() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def _run_resblock1(x, w1, b1, w2, b2, d1):
    """hg_op_resblock1: x [B,C,L] cpu -> (y [B,C,L], fused?)"""
    L = _native.lib()
    dev = torch.device("cuda", 0)
    B, C, n = x.shape
    k, np_ = w1[0].shape[-1], len(w1)
    xc = x.transpose(1, 2).contiguous().to(dev)
    y = torch.full((B, n, C), float("nan"), device=dev)
    arr = lambda ts: (ctypes.c_void_p * np_)(*[t.data_ptr() for t in ts])  # noqa: E731
    w1c, b1c, w2c, b2c = ([t.contiguous() for t in ts] for ts in (w1, b1, w2, b2))
    fused = ctypes.c_int32(-1)
    _native.check(L.hg_op_resblock1(0, xc.data_ptr(), B, n, C, k, np_, (ctypes.c_int32 * np_)(*d1), arr(w1c), arr(b1c), arr(w2c),
                                    arr(b2c), 0.1, y.data_ptr(), torch.cuda.current_stream().cuda_stream, ctypes.byref(fused)))
    torch.cuda.synchronize()
    return y, fused.value


@pytest.mark.parametrize("C", [64, 32])
@pytest.mark.parametrize("n", [1, 100, 232, 233, 464, 488, 489, 1000, 4104])
def test_fused_resblock_parity(C, n):
    """The fused-ResBlock kernel (conv_chain_tc.cu: the three (c1, c2) pairs of a k = 3 ResBlock1 in one launch,
    residual stream and operand kept on chip) against the pair-by-pair schedule — bit for bit, every operation
    happens in the same order — and against the fp64 restatement of hifi/models.py:88-95 with the kernels' operand
    model.  Lengths straddle the 232 / 488-row output tiles."""
    import os
    g = torch.Generator().manual_seed(C * 100 + n)
    k, d1 = 3, (1, 3, 5)
    x = torch.randn(2, C, n, generator=g)
    w1 = [torch.randn(C, C, k, generator=g) / (C * k) ** 0.5 for _ in d1]
    w2 = [torch.randn(C, C, k, generator=g) / (C * k) ** 0.5 for _ in d1]
    b1 = [torch.randn(C, generator=g) * 0.1 for _ in d1]
    b2 = [torch.randn(C, generator=g) * 0.1 for _ in d1]
    os.environ["HG_CHAIN"] = "1"  # the fused kernel is off by default (slower than the three pairs, see api.cu)
    try:
        y, fused = _run_resblock1(x, w1, b1, w2, b2, d1)
    finally:
        del os.environ["HG_CHAIN"]
    assert fused == 1
    y_pairs, fused0 = _run_resblock1(x, w1, b1, w2, b2, d1)
    assert fused0 == 0
    assert not torch.isnan(y).any()
    assert torch.equal(y, y_pairs)
    r = x.double()
    for m, d in enumerate(d1):
        a = bf16_round(F.leaky_relu(r.float(), 0.1))
        xt = F.conv1d(a, bf16_round(w1[m]), b1[m].double(), dilation=d, padding=d)
        xt = bf16_round(F.leaky_relu(xt.float(), 0.1))
        r = (F.conv1d(xt, bf16_round(w2[m]), b2[m].double(), padding=1) + r).float().double()  # the stream is fp32
    err = (y.cpu().transpose(1, 2).double() - r).abs().max().item()
    assert err <= 4e-3 * r.abs().max().item(), err


def test_tcgen05_descriptor_selftest():
    n, report = _native.selftest(0)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    open("gpurun_out/selftest_tcgen05.txt", "w").write(report)
    mode = int(os.environ.get("HG_DESC_MODE", "0"))
    bad = [ln for ln in report.splitlines() if f"mode={mode} " in ln and "MISMATCH" in ln]
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("prec", ["fp32_ffma", "fp32", "bf16"])
@pytest.mark.parametrize("n", [1, 7, 129, 800])
def test_conv_pre_padded_k(n, prec):
    """conv_pre's shape (hifi/models.py:152-154: Conv1d(80, 512, 7, padding=3)) on its own: 80 input channels
    are not a multiple of the 64-wide K chunk, so the tensor-core path zero-pads K to 128 in the operand
    and in the packed weights.  No activation on the input (slope 1.0), as in Generator.forward :186."""
    g = torch.Generator().manual_seed(80 + n)
    x = torch.randn(2, 80, n, generator=g)
    w = torch.randn(512, 80, 7, generator=g) / (80 * 7) ** 0.5
    b = torch.randn(512, generator=g) * 0.1
    y = run_conv1d(x, w, b, 1, 1.0, None, prec)
    ref = ref_conv1d(x, w, b, 1, 1.0, None, prec)
    assert not torch.isnan(y).any()
    assert (y.double() - ref).abs().max().item() <= TOL[prec] * max(1.0, ref.abs().max().item())
