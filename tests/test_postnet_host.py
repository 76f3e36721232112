"""N3: the PostNet / mel_linear oracle against the reference's own outputs
(tests/golden/postnet.npz, tools/make_golden.py), the mirror module's state_dict schema and seeded
initialisation, and the BatchNorm fold.  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fixtures as fx
from oracle import postnet_oracle as po
from oracle.common import max_abs
from tts_king_b200.fs_two.transformer.Layers import PostNet

from _util import golden


def _float_state(m):
    return {k: v for k, v in m.state_dict().items() if v.dtype.is_floating_point}


def full_alive_state():
    """The full-size PostNet's seeded weights, rebuilt through the package's own constructor."""
    g = golden("postnet")
    torch.manual_seed(1234)
    m = PostNet(**fx.POSTNET_FULL)
    assert fx.state_digest(_float_state(m)) == str(g["full.digest_fresh"])  # same init as the reference's class
    sd = fx.alive_batchnorm_({k: v.clone() for k, v in m.state_dict().items()})
    assert fx.state_digest({k: v for k, v in sd.items() if v.dtype.is_floating_point}) == str(g["full.digest_alive"])
    return sd


def test_state_dict_schema():
    m = PostNet()
    sd = m.state_dict()
    assert len(sd) == 35
    assert list(sd.keys())[:7] == ["convolutions.0.0.conv.weight", "convolutions.0.0.conv.bias", "convolutions.0.1.weight",
                                   "convolutions.0.1.bias", "convolutions.0.1.running_mean", "convolutions.0.1.running_var",
                                   "convolutions.0.1.num_batches_tracked"]
    assert tuple(sd["convolutions.0.0.conv.weight"].shape) == (512, 80, 5)
    assert tuple(sd["convolutions.2.0.conv.weight"].shape) == (512, 512, 5)
    assert tuple(sd["convolutions.4.0.conv.weight"].shape) == (80, 512, 5)
    assert tuple(sd["convolutions.4.1.running_var"].shape) == (80,)


def test_oracle_matches_reference_postnet_tiny():
    g = golden("postnet")
    sd = {k[len("tiny.sd."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("tiny.sd.")}
    y = po.postnet_forward(sd, torch.from_numpy(g["tiny.x"]))
    assert max_abs(y.numpy(), g["tiny.y"]) <= 2e-6
    m = PostNet(**fx.POSTNET_TINY)
    m.load_state_dict(sd)  # the reference's keys load strictly
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "running_var" not in k})


def test_oracle_matches_reference_postnet_full_and_mel_tail():
    g = golden("postnet")
    sd = full_alive_state()
    out, post_out = po.mel_tail(torch.from_numpy(g["full.lin_w"]), torch.from_numpy(g["full.lin_b"]), sd,
                                torch.from_numpy(g["full.decoder_output"]))
    assert max_abs(out.numpy(), g["full.output"]) <= 2e-6
    assert max_abs(post_out.numpy(), g["full.postnet_output"]) <= 2e-5
    assert max_abs(po.postnet_forward(sd, torch.from_numpy(g["full.output"])).numpy(), g["full.postnet"]) <= 2e-5
    assert max_abs(po.postnet_forward(sd, torch.from_numpy(g["full.x_T1"])).numpy(), g["full.y_T1"]) <= 2e-5
    # the golden signal is alive at every depth: not a fixture that would pass with a broken fold
    assert np.abs(g["full.postnet"]).max() > 0.5 and np.abs(g["full.postnet"]).std() > 0.1


def test_batchnorm_fold_is_the_eval_mode_map():
    g = golden("postnet")
    sd = {k[len("tiny.sd."):]: torch.from_numpy(g[k]).double() for k in g.files if k.startswith("tiny.sd.")}
    m = PostNet(**fx.POSTNET_TINY)
    m.load_state_dict({k: v.float() if v.dtype.is_floating_point else v for k, v in sd.items()})
    h = torch.from_numpy(g["tiny.x"]).double().transpose(1, 2)
    for i, (w, b) in enumerate(m.folded_layers()):
        wo, bo = po.fold_batchnorm(sd, i)
        assert max_abs(w.numpy(), wo.numpy()) <= 1e-6 and max_abs(b.numpy(), bo.numpy()) <= 1e-6
        h = F.conv1d(h, wo, bo, padding=2)
        if i < 4:
            h = torch.tanh(h)
    assert max_abs(h.transpose(1, 2).numpy(), po.postnet_forward(sd, torch.from_numpy(g["tiny.x"]).double()).numpy()) <= 1e-12


def test_cpu_and_training_mode_are_refused():
    m = PostNet(**fx.POSTNET_TINY)
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.zeros(1, 4, 80))
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, 80))
