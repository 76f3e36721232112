"""Deterministic weights / inputs shared by tools/make_golden.py, tests/ and bench.py.

TEST INFRASTRUCTURE ONLY.  Input recipes are SURVEY.md §8(d)'s: seed 1234 for weights (the
reference's own ``hifi.seed``, ``config.yaml:23``), seed 7 for mels.
"""
from __future__ import annotations

import hashlib
from typing import Dict

import numpy as np
import torch

from .common import GenConfig, conv_specs


class AttrDict(dict):
    """Attribute-and-item access, the shape of object the reference passes as ``h``."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def make_h(cfg: GenConfig) -> AttrDict:
    return AttrDict(
        resblock=cfg.resblock,
        upsample_rates=list(cfg.upsample_rates),
        upsample_kernel_sizes=list(cfg.upsample_kernel_sizes),
        upsample_initial_channel=cfg.upsample_initial_channel,
        resblock_kernel_sizes=list(cfg.resblock_kernel_sizes),
        resblock_dilation_sizes=[list(d) for d in cfg.resblock_dilation_sizes],
        MAX_WAV_VALUE=32768,
    )


# configurations exercised by the fixtures
V1 = GenConfig(512, (8, 8, 2, 2), (16, 16, 4, 4), (3, 7, 11), ((1, 3, 5),) * 3, "1")
V2_NARROW = GenConfig(128, (8, 8, 2, 2), (16, 16, 4, 4), (3, 7, 11), ((1, 3, 5),) * 3, "1")
V3_RB2 = GenConfig(256, (8, 8, 4), (16, 16, 8), (3, 5, 7), ((1, 2), (2, 6), (3, 12)), "2")
TINY_RB1 = GenConfig(32, (8, 8, 2, 2), (16, 16, 4, 4), (3, 7, 11), ((1, 3, 5),) * 3, "1")
TINY_RB2 = GenConfig(32, (8, 8, 4), (16, 16, 8), (3, 5, 7), ((1, 2), (2, 6), (3, 12)), "2")


def synthetic_mel(B: int, T: int, seed: int = 7, kind: str = "randn") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 80, T, generator=g)
    if kind == "logmel":  # range implied by hifi/meldataset.py: log(clamp(x, 1e-5))
        x = torch.clamp(x * 2.0 - 5.0, -11.5129, 2.0)
    return x


def alive_state(cfg: GenConfig, seed: int = 4321, gain: float = 0.7) -> Dict[str, torch.Tensor]:
    """Folded-layout state dict with trained-like gains: weights
    N(0, gain^2/fan_in), biases N(0, 0.05^2).  Default random init attenuates the signal ~30x by
    the last stage; these keep every layer 'alive' so rounding errors are not hidden."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for s in conv_specs(cfg):
        fan_in = s.cin * s.k / (s.stride if s.kind == "convT" else 1)
        sd[f"{s.name}.weight"] = torch.randn(*s.weight_shape, generator=g) * (gain / fan_in ** 0.5)
        sd[f"{s.name}.bias"] = torch.randn(s.cout, generator=g) * 0.05
    return sd


def state_digest(state: Dict[str, torch.Tensor]) -> str:
    """sha256 over (key, fp32 bytes) in sorted-key order."""
    h = hashlib.sha256()
    for k in sorted(state.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(state[k].detach().cpu().numpy().astype(np.float32)).tobytes())
    return h.hexdigest()


def tensor_digest(t: torch.Tensor) -> str:
    """sha256 of one tensor's fp32 bytes (seeded inputs are regenerated in the tests, not stored)."""
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32)).tobytes()).hexdigest()


def to_numpy_state(state: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy() for k, v in state.items()}


# ----------------------------------------------------------------------------- PostNet (N3)
POSTNET_FULL = dict(n_mel_channels=80, postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5)
POSTNET_TINY = dict(n_mel_channels=80, postnet_embedding_dim=32, postnet_kernel_size=5, postnet_n_convolutions=5)


def alive_batchnorm_(state: Dict[str, torch.Tensor], seed: int = 97) -> Dict[str, torch.Tensor]:
    """Give every BatchNorm of a PostNet state dict trained-like statistics, in place and seeded: a
    fresh BatchNorm1d (mean 0, var 1, gamma 1, beta 0) would make the eval-mode fold a no-op and
    hide mistakes in it.  Conv weights are rescaled so activations stay O(1) through the tanh's."""
    g = torch.Generator().manual_seed(seed)
    i = 0
    while f"convolutions.{i}.1.running_mean" in state:
        c = state[f"convolutions.{i}.1.running_mean"].numel()
        state[f"convolutions.{i}.1.running_mean"] = torch.randn(c, generator=g) * 0.2
        state[f"convolutions.{i}.1.running_var"] = torch.rand(c, generator=g) + 0.5
        state[f"convolutions.{i}.1.weight"] = torch.rand(c, generator=g) + 0.5
        state[f"convolutions.{i}.1.bias"] = torch.randn(c, generator=g) * 0.1
        w = state[f"convolutions.{i}.0.conv.weight"]
        state[f"convolutions.{i}.0.conv.weight"] = w * (1.5 / (w.std() * (w.shape[1] * w.shape[2]) ** 0.5))
        i += 1
    return state
