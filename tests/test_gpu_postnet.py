"""N3: PostNet + mel_linear on the GPU against the reference's own outputs
(tests/golden/postnet.npz) and the time-major hand-off into the vocoder.  Needs a B200.

Tolerances: fp32 paths max-abs <= 1e-4 relative to the signal's peak (the PostNet's output is a
log-mel correction of O(1), not a waveform in [-1, 1]); bf16 path SNR >= 40 dB."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import postnet_oracle as po
from oracle.common import max_abs, snr_db
from tts_king_b200.fs_two.model.fastspeech2 import MelLinear, mel_tail, mel_to_vocoder
from tts_king_b200.fs_two.transformer.Layers import PostNet

from _util import golden, make_generator
from test_postnet_host import full_alive_state

pytestmark = pytest.mark.gpu

FP32_REL_TOL = 1e-4
BF16_SNR_DB = 40.0


def tiny_postnet(prec):
    g = golden("postnet")
    sd = {k[len("tiny.sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("tiny.sd.")}
    m = PostNet(**fx.POSTNET_TINY, precision=prec)
    m.load_state_dict(sd)
    return m.eval().cuda(), g


@pytest.mark.parametrize("prec", ["fp32", "fp32_ffma"])
def test_postnet_tiny_fp32_matches_reference(prec):
    m, g = tiny_postnet(prec)
    y = m(torch.from_numpy(g["tiny.x"]).cuda()).cpu().numpy()
    assert y.shape == g["tiny.y"].shape
    assert max_abs(y, g["tiny.y"]) <= FP32_REL_TOL * np.abs(g["tiny.y"]).max()


def test_postnet_tiny_bf16_snr():
    m, g = tiny_postnet("bf16")
    y = m(torch.from_numpy(g["tiny.x"]).cuda()).cpu().numpy()
    assert snr_db(g["tiny.y"], y) >= BF16_SNR_DB


@pytest.mark.parametrize("prec", ["fp32", "fp32_ffma", "bf16"])
def test_mel_tail_full_size_matches_reference(prec):
    """fastspeech2.py:101-104 at the reference's sizes (256 -> 80, then 80 -> 512 x3 -> 80)."""
    g = golden("postnet")
    post = PostNet(**fx.POSTNET_FULL, precision=prec)
    post.load_state_dict(full_alive_state())
    post.eval().cuda()
    lin = MelLinear(256, 80, precision=prec)
    lin.load_state_dict({"weight": torch.from_numpy(g["full.lin_w"]), "bias": torch.from_numpy(g["full.lin_b"])})
    lin.cuda()
    output, postnet_output = mel_tail(torch.from_numpy(g["full.decoder_output"]).cuda(), lin, post)
    assert output.shape == postnet_output.shape == (3, 41, 80) and postnet_output.is_contiguous()
    if prec == "bf16":
        assert snr_db(g["full.output"], output.cpu().numpy()) >= BF16_SNR_DB
        assert snr_db(g["full.postnet_output"], postnet_output.cpu().numpy()) >= BF16_SNR_DB
        return
    assert max_abs(output.cpu().numpy(), g["full.output"]) <= FP32_REL_TOL * np.abs(g["full.output"]).max()
    assert max_abs(postnet_output.cpu().numpy(), g["full.postnet_output"]) <= FP32_REL_TOL * np.abs(g["full.postnet_output"]).max()
    # PostNet.forward alone (no residual), T = 1, and a non-contiguous input view
    y = post(torch.from_numpy(g["full.output"]).cuda())
    assert max_abs(y.cpu().numpy(), g["full.postnet"]) <= FP32_REL_TOL * np.abs(g["full.postnet"]).max()
    y1 = post(torch.from_numpy(g["full.x_T1"]).cuda())
    assert max_abs(y1.cpu().numpy(), g["full.y_T1"]) <= FP32_REL_TOL * max(1.0, np.abs(g["full.y_T1"]).max())
    xt = torch.from_numpy(g["full.output"]).cuda().transpose(1, 2).contiguous().transpose(1, 2)  # [B,T,80] view of [B,80,T]
    assert not xt.is_contiguous()
    assert torch.equal(post(xt), y)
    assert torch.equal(post.forward_residual(xt), post.forward_residual(xt.contiguous()))


def test_postnet_against_oracle_on_a_longer_batch():
    """Sizes the golden file does not hold: the CUDA stack against the CPU oracle on seeded inputs."""
    sd = full_alive_state()
    post = PostNet(**fx.POSTNET_FULL)
    post.load_state_dict(sd)
    post.eval().cuda()
    x = torch.randn(4, 333, 80, generator=torch.Generator().manual_seed(12))
    ref = po.postnet_forward({k: v.double() for k, v in sd.items()}, x.double()).numpy()
    y = post(x.cuda()).cpu().numpy()
    assert max_abs(y, ref) <= FP32_REL_TOL * np.abs(ref).max()


def test_time_major_hand_off_into_the_vocoder():
    """tts_king.py:48: the vocoder reads the PostNet's [B,T,80] output through a transposed view."""
    gen = make_generator(fx.V1).cuda()
    mel_tm = torch.randn(2, 40, 80, generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        a = gen(mel_to_vocoder(mel_tm))
        b = gen(mel_tm.transpose(1, 2).contiguous())
    assert torch.equal(a, b)


def test_refusals_on_gpu():
    m, g = tiny_postnet("fp32")
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 5, 81).cuda())
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 5, 80))  # CPU input, CUDA weights
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(2, 5, 80).cuda())
