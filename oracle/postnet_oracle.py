"""CPU restatement of the step before the vocoder — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the
product path (tts_king_b200/) never does.

Restates, with plain torch functional calls on explicit state dicts:
  * PostNet.forward            reference fs_two/transformer/Layers.py:133-143 (eval mode)
  * mel_linear + postnet + add reference fs_two/model/fastspeech2.py:101-104
Pinned against outputs of the reference's own PostNet run in this container
(tools/make_golden.py -> tests/golden/postnet.npz; tests/test_oracle.py).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm1d default, Layers.py:97,114,129


def n_convolutions(state: Dict[str, torch.Tensor]) -> int:
    n = 0
    while f"convolutions.{n}.0.conv.weight" in state:
        n += 1
    return n


def postnet_forward(state: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """x [B,T,n_mel] -> [B,T,n_mel].  Layers.py:133-143 with self.training == False: the dropouts are
    identities and BatchNorm1d normalises with its running statistics."""
    n = n_convolutions(state)
    h = x.contiguous().transpose(1, 2)                                   # :134
    for i in range(n):
        w, b = state[f"convolutions.{i}.0.conv.weight"], state[f"convolutions.{i}.0.conv.bias"]
        h = F.conv1d(h, w, b, padding=(w.shape[-1] - 1) // 2)            # ConvNorm, :64-67
        mean, var = state[f"convolutions.{i}.1.running_mean"], state[f"convolutions.{i}.1.running_var"]
        g, beta = state[f"convolutions.{i}.1.weight"], state[f"convolutions.{i}.1.bias"]
        h = (h - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS) * g[None, :, None] + beta[None, :, None]
        if i < n - 1:
            h = torch.tanh(h)                                            # :137-139
    return h.contiguous().transpose(1, 2)                                # :142


def fold_batchnorm(state: Dict[str, torch.Tensor], i: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv weight/bias of layer i with its eval-mode BatchNorm folded in (an affine map per output
    channel): w' = w * s, b' = (b - mean) * s + beta, s = gamma / sqrt(var + eps)."""
    w, b = state[f"convolutions.{i}.0.conv.weight"], state[f"convolutions.{i}.0.conv.bias"]
    s = state[f"convolutions.{i}.1.weight"] / torch.sqrt(state[f"convolutions.{i}.1.running_var"] + BN_EPS)
    return w * s[:, None, None], (b - state[f"convolutions.{i}.1.running_mean"]) * s + state[f"convolutions.{i}.1.bias"]


def mel_tail(lin_w: torch.Tensor, lin_b: torch.Tensor, state: Dict[str, torch.Tensor],
             decoder_output: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fastspeech2.py:101-104: output = mel_linear(decoder_output); postnet_output = postnet(output) + output."""
    output = F.linear(decoder_output, lin_w, lin_b)
    return output, postnet_forward(state, output) + output
