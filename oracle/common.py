"""Shared helpers for the CPU oracles (TEST INFRASTRUCTURE ONLY).

Config handling mirrors what ``Generator.__init__`` reads from ``h``
(reference ``hifi/models.py:147-183``); the conv enumeration order is the module
registration order there, which is also the state_dict order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np

NUM_MELS = 80  # hard-coded at hifi/models.py:153


@dataclass(frozen=True)
class ConvSpec:
    name: str          # state_dict prefix, e.g. "resblocks.4.convs1.2"
    kind: str          # "conv" | "convT"
    cin: int
    cout: int
    k: int
    dilation: int = 1  # conv only
    stride: int = 1    # convT only

    @property
    def weight_shape(self):
        # Conv1d [C_out, C_in, k]; ConvTranspose1d [C_in, C_out, k]
        if self.kind == "conv":
            return (self.cout, self.cin, self.k)
        return (self.cin, self.cout, self.k)


@dataclass(frozen=True)
class GenConfig:
    upsample_initial_channel: int
    upsample_rates: Sequence[int]
    upsample_kernel_sizes: Sequence[int]
    resblock_kernel_sizes: Sequence[int]
    resblock_dilation_sizes: Sequence[Sequence[int]]
    resblock: str = "1"

    @property
    def hop(self) -> int:
        return int(np.prod(self.upsample_rates))


def config_from_h(h) -> GenConfig:
    """Accepts AttrDict / OmegaConf / SimpleNamespace / dict, like the reference."""
    def get(k):
        try:
            return getattr(h, k)
        except AttributeError:
            return h[k]

    return GenConfig(
        upsample_initial_channel=int(get("upsample_initial_channel")),
        upsample_rates=tuple(int(v) for v in get("upsample_rates")),
        upsample_kernel_sizes=tuple(int(v) for v in get("upsample_kernel_sizes")),
        resblock_kernel_sizes=tuple(int(v) for v in get("resblock_kernel_sizes")),
        resblock_dilation_sizes=tuple(tuple(int(d) for d in ds) for ds in get("resblock_dilation_sizes")),
        resblock=str(get("resblock")),
    )


V1 = GenConfig(512, (8, 8, 2, 2), (16, 16, 4, 4), (3, 7, 11), ((1, 3, 5),) * 3, "1")


def conv_specs(cfg: GenConfig) -> List[ConvSpec]:
    """Every conv of the generator in module order (hifi/models.py:152-181)."""
    uic = cfg.upsample_initial_channel
    specs = [ConvSpec("conv_pre", "conv", NUM_MELS, uic, 7)]
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        specs.append(ConvSpec(f"ups.{i}", "convT", uic >> i, uic >> (i + 1), k, stride=u))
    ch = uic
    nd = 3 if cfg.resblock == "1" else 2  # ResBlock1 reads dilation[0..2], ResBlock2 dilation[0..1]
    for i in range(len(cfg.upsample_rates)):
        ch = uic >> (i + 1)
        for j, (k, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            n = i * len(cfg.resblock_kernel_sizes) + j
            if cfg.resblock == "1":
                for m, d in enumerate(ds[:nd]):
                    specs.append(ConvSpec(f"resblocks.{n}.convs1.{m}", "conv", ch, ch, k, dilation=d))
                for m, _ in enumerate(ds[:nd]):
                    specs.append(ConvSpec(f"resblocks.{n}.convs2.{m}", "conv", ch, ch, k, dilation=1))
            else:
                for m, d in enumerate(ds[:nd]):
                    specs.append(ConvSpec(f"resblocks.{n}.convs.{m}", "conv", ch, ch, k, dilation=d))
    specs.append(ConvSpec("conv_post", "conv", ch, 1, 7))
    return specs


def fold_weight_norm(v: np.ndarray, g: np.ndarray) -> np.ndarray:
    """w = v * g / ||v|| with the norm over all dims but 0 (torch._weight_norm(v,g,0));
    reached from remove_weight_norm, hifi/models.py:97-101,203-210."""
    v64 = v.astype(np.float64)
    norm = np.sqrt((v64 * v64).reshape(v.shape[0], -1).sum(axis=1)).reshape(-1, *([1] * (v.ndim - 1)))
    return (v64 * (g.astype(np.float64).reshape(norm.shape) / norm)).astype(v.dtype)


def folded_weights(state: Dict[str, np.ndarray], cfg: GenConfig) -> List[np.ndarray]:
    """state_dict (g/v layout or folded layout) -> flat [w, b, w, b, ...] in module order."""
    out: List[np.ndarray] = []
    for s in conv_specs(cfg):
        if f"{s.name}.weight" in state:
            w = np.asarray(state[f"{s.name}.weight"])
        else:
            w = fold_weight_norm(np.asarray(state[f"{s.name}.weight_v"]), np.asarray(state[f"{s.name}.weight_g"]))
        assert tuple(w.shape) == s.weight_shape, (s.name, w.shape, s.weight_shape)
        out.append(np.ascontiguousarray(w))
        out.append(np.ascontiguousarray(np.asarray(state[f"{s.name}.bias"])))
    return out


# ---- metrics used by the parity tests -------------------------------------------------

def max_abs(a, b) -> float:
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def snr_db(ref, test) -> float:
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    return float(10.0 * np.log10(np.sum(ref * ref) / max(np.sum(err * err), 1e-300)))


def ac_snr_db(ref, test) -> float:
    """SNR with the reference mean removed from the signal power (random-init output is
    DC-dominated)."""
    ref = np.asarray(ref, dtype=np.float64)
    err = np.asarray(test, dtype=np.float64) - ref
    ac = ref - ref.mean()
    return float(10.0 * np.log10(np.sum(ac * ac) / max(np.sum(err * err), 1e-300)))
