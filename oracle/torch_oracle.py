"""``torch.nn.functional`` restatement of the generator on CPU tensors.

TEST INFRASTRUCTURE ONLY.  This is the "port" CPU baseline bench.py times: it issues the
same ATen convolution / elementwise calls the reference's eager modules dispatch
(78 conv, 77 leaky_relu, 36 add, 8 add_, 4 div, 1 tanh), on folded weights,
under ``torch.no_grad`` — i.e. what ``HIFIapi.generate`` runs (reference ``hifiapi.py:47-49``).

Follows reference ``hifi/models.py:185-201`` (Generator.forward), ``:88-95``
(ResBlock1.forward), ``:134-139`` (ResBlock2.forward) and ``hifi/vocoder/utils.py:36-37``.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .common import GenConfig

LRELU_SLOPE = 0.1  # hifi/models.py:9


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return int((kernel_size * dilation - dilation) / 2)


def fold_state_dict(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """g/v layout -> folded layout via torch._weight_norm(v, g, 0), the function
    remove_weight_norm bakes in (hifi/models.py:203-210)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in state.items():
        if k.endswith(".weight_g"):
            continue
        if k.endswith(".weight_v"):
            g = state[k[: -len("weight_v")] + "weight_g"]
            out[k[: -len("_v")]] = torch._weight_norm(v, g, 0)
        else:
            out[k] = v
    return out


@torch.no_grad()
def forward(cfg: GenConfig, state: Dict[str, torch.Tensor], mel: torch.Tensor) -> torch.Tensor:
    """mel [B,80,T] (any strides, CPU) -> wav [B,1,T*hop]; dtype follows ``mel``."""
    sd = {k: v.to(mel.dtype) for k, v in fold_state_dict(state).items()}
    nk = len(cfg.resblock_kernel_sizes)
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            p = f"resblocks.{i * nk + j}"
            r = x
            for m, d in enumerate(ds[:3 if cfg.resblock == "1" else 2]):
                xt = F.leaky_relu(r, LRELU_SLOPE)
                if cfg.resblock == "1":
                    xt = F.conv1d(xt, sd[f"{p}.convs1.{m}.weight"], sd[f"{p}.convs1.{m}.bias"],
                                  dilation=d, padding=get_padding(rk, d))
                    xt = F.leaky_relu(xt, LRELU_SLOPE)
                    xt = F.conv1d(xt, sd[f"{p}.convs2.{m}.weight"], sd[f"{p}.convs2.{m}.bias"],
                                  padding=get_padding(rk, 1))
                else:
                    xt = F.conv1d(xt, sd[f"{p}.convs.{m}.weight"], sd[f"{p}.convs.{m}.bias"],
                                  dilation=d, padding=get_padding(rk, d))
                r = xt + r
            xs = r if xs is None else xs.add_(r)
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01, hifi/models.py:197
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)
