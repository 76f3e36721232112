"""Parity at the sizes BASELINE.json names (cfg-1 1x256, cfg-2 16x800, cfg-3 64x2000, cfg-5 one long mel),
against outputs of the reference itself where the reference can produce them in seconds
(tests/golden/v1_baseline_shapes.npz, tools/make_golden.py) and through size-independent properties
(batch independence, chunk == monolithic) where it cannot.  Needs a B200: `-m gpu` under gpurun.
"""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import torch_oracle
from oracle.common import max_abs, snr_db
from tts_king_b200 import parallel

from _util import golden, make_generator

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # north_star: fp32 max-abs error <= 1e-4
BF16_SNR_DB = 40.0   # north_star: bf16 path SNR >= 40 dB against the fp32 reference


def _fwd(m, mel):
    with torch.no_grad():
        y = m(mel.cuda())
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("tag", ["cfg1", "utt800"])
def test_reference_outputs_at_baseline_shapes(tag):
    """cfg-1 (1 x 80 x 256, seed 7) and one 800-frame utterance of cfg-2 against the reference Generator's own
    waveform: fp32 path <= 1e-4, bf16 path >= 40 dB, graph replay bit-identical to eager."""
    g = golden("v1_baseline_shapes")
    frames = int(g[f"{tag}.frames"])
    mel = fx.synthetic_mel(1, frames, seed=7)
    assert fx.tensor_digest(mel) == str(g[f"{tag}.mel_sha256"])
    ref = g[f"{tag}.y"]
    m = make_generator(fx.V1, precision="fp32")
    assert fx.state_digest(m.state_dict()) == str(g["digest_folded"])
    m.cuda()
    y32 = _fwd(m, mel)
    assert y32.shape == ref.shape
    assert max_abs(y32.cpu().numpy(), ref) <= FP32_TOL
    run = m.make_graphed(1, frames)
    assert torch.equal(run(mel.cuda()), y32)
    m.precision = "bf16"
    y16 = _fwd(m, mel).cpu().numpy()
    assert snr_db(ref, y16) >= BF16_SNR_DB, snr_db(ref, y16)


def test_cfg2_batch_against_reference_utterance():
    """cfg-2 (16 x 800): item 0 of the batch is the golden utterance; inside a batch of 16 it must give
    the same samples as alone (bitwise), hence the reference's waveform within tolerance."""
    g = golden("v1_baseline_shapes")
    mel = fx.synthetic_mel(16, 800, seed=70)
    mel[0] = fx.synthetic_mel(1, 800, seed=7)[0]
    ref = g["utt800.y"]
    for prec in ("fp32", "bf16"):
        m = make_generator(fx.V1, precision=prec).cuda()
        yb = _fwd(m, mel)
        y0 = _fwd(m, mel[:1])
        assert torch.equal(yb[:1], y0), prec
        if prec == "fp32":
            assert max_abs(yb[:1].cpu().numpy(), ref) <= FP32_TOL
        else:
            assert snr_db(ref, yb[:1].cpu().numpy()) >= BF16_SNR_DB
        del yb, y0, m
        torch.cuda.empty_cache()


def test_cfg3_size_batch_is_exact_per_item():
    """cfg-3 (64 x 2000 frames, bf16): three sampled items are bit-equal to their own B = 1 forward, and
    one of them matches the CPU oracle (the reference's arithmetic) on all 2000 frames."""
    B, T = 64, 2000
    mel = fx.synthetic_mel(B, T, seed=33)
    m = make_generator(fx.V1, precision="bf16").cuda()
    y = _fwd(m, mel)
    assert y.shape == (B, 1, T * 256)
    assert torch.isfinite(y).all()
    for i in (0, 31, 63):
        yi = _fwd(m, mel[i:i + 1])
        assert torch.equal(y[i:i + 1], yi), i
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = torch_oracle.forward(fx.V1, sd, mel[31:32]).numpy()
    assert snr_db(ref, y[31:32].cpu().numpy()) >= BF16_SNR_DB
    del y
    torch.cuda.empty_cache()
    m.precision = "fp32"
    y32 = _fwd(m, mel[31:32]).cpu().numpy()
    assert max_abs(y32, ref) <= FP32_TOL


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_cfg5_shaped_long_form_chunks(prec):
    """cfg-5's mechanism at a length the test can afford (1 x 41 000 frames = 8 minutes of audio; the
    60-minute mel is 310 078): chunked_forward_into (16 384-frame chunks, 13-frame halo, conv_post stores
    each chunk's samples in place) equals the monolithic forward on windows that straddle every chunk
    boundary and on both ends — bit for bit — and int16 output agrees with the float one."""
    T = 41000
    m = make_generator(fx.V1, precision=prec).cuda()
    halo, hop = m.halo_frames, m.hop_length
    assert halo == 13 and hop == 256
    mel = fx.synthetic_mel(1, T, seed=55).cuda()
    out = torch.full((1, 1, T * hop), float("nan"), device="cuda")
    with torch.no_grad():
        parallel.chunked_forward_into(m, mel, 16384, halo, out)
        torch.cuda.synchronize()
        assert torch.isfinite(out).all()
        chunks = parallel.plan_time_chunks(T, 3, halo)
        assert [c.frames for c in chunks] == [13667, 13667, 13666]
        edges = [0] + [c.stop for c in chunks[:-1]] + [T]
        for e in edges:
            a, b = max(0, e - 300), min(T, e + 300)       # window owned across the boundary
            lo, hi = max(0, a - halo), min(T, b + halo)   # plus its receptive halo
            w = m(mel[:, :, lo:hi])
            assert torch.equal(out[:, :, a * hop:b * hop], w[:, :, (a - lo) * hop:(b - lo) * hop]), e
        o16 = torch.zeros((1, 1, T * hop), dtype=torch.int16, device="cuda")
        parallel.chunked_forward_into(m, mel, 16384, halo, o16)
        want = (out * 32768.0).to(torch.int32).to(torch.int16)  # truncation toward zero, low 16 bits
        assert torch.equal(o16, want)


def test_int16_tail_matches_numpy_cast_including_the_wrap():
    """HIFIapi.generate's tail (hifiapi.py:50-51: audio * 32768 -> .numpy().astype('int16')) fused into
    conv_post.  (a) exactly +1.0 wraps to -32768 and exactly -1.0 gives -32768, as numpy does on the
    reference's platform (tests/golden/int16_cast.npz, produced by tools/make_golden.py); (b) on a
    full-scale signal the fused cast equals numpy's cast of the float output bit for bit."""
    cast = golden("int16_cast")
    table = {float(x): int(y) for x, y in zip(cast["x"], cast["y"])}
    assert table[1.0] == -32768 and table[-1.0] == -32768 and table[0.0] == 0
    g = golden("tiny_rb1")
    mel = torch.from_numpy(g["mel"]).cuda()
    for bias, want in ((30.0, table[1.0]), (-30.0, table[-1.0]), (0.0, table[0.0])):
        m = make_generator(fx.TINY_RB1, seed=6, fold=True)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["conv_post.weight"].zero_()
        sd["conv_post.bias"].fill_(bias)       # tanh(+-30) == +-1.0 exactly in fp32
        m.load_state_dict(sd)
        m.cuda()
        with torch.no_grad():
            yf = m(mel)
            yi = m.generate_int16(mel)
        assert (yf == (1.0 if bias > 0 else -1.0 if bias < 0 else 0.0)).all()
        assert yi.dtype == torch.int16 and (yi == want).all(), (bias, int(yi.flatten()[0]))
    m = make_generator(fx.TINY_RB1, seed=6, fold=True)
    sd = {k: torch.from_numpy(g[k]) for k in g.files if k.startswith("alive.")}
    m.load_state_dict({k[len("alive."):]: v for k, v in sd.items()})
    m.cuda()
    with torch.no_grad():
        yf = m(mel).cpu().numpy()
        yi = m.generate_int16(mel).cpu().numpy()
    assert np.abs(yf).max() > 0.5  # full-scale: both signs, many fractional parts
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = (yf * np.float32(32768.0)).astype("int16")
    assert np.array_equal(yi, want)
