"""ctypes front end of ``hifigan_oracle.c`` (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from .common import GenConfig, conv_specs, folded_weights

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhifigan_oracle.so")
_lib: Optional[ctypes.CDLL] = None

OG_MAX_UPS, OG_MAX_KERNELS, OG_MAX_DIL = 8, 8, 4


class _OgConfig(ctypes.Structure):
    _fields_ = [
        ("num_mels", ctypes.c_int),
        ("upsample_initial_channel", ctypes.c_int),
        ("num_upsamples", ctypes.c_int),
        ("upsample_rates", ctypes.c_int * OG_MAX_UPS),
        ("upsample_kernel_sizes", ctypes.c_int * OG_MAX_UPS),
        ("num_kernels", ctypes.c_int),
        ("resblock_kernel_sizes", ctypes.c_int * OG_MAX_KERNELS),
        ("num_dilations", ctypes.c_int),
        ("resblock_dilation_sizes", (ctypes.c_int * OG_MAX_DIL) * OG_MAX_KERNELS),
        ("resblock_type", ctypes.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile the C oracle in-tree (gcc, OpenMP).  Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("hifigan_oracle.c", "hifigan_oracle_impl.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.og_forward_f32.restype = ctypes.c_int
        _lib.og_forward_f64.restype = ctypes.c_int
        _lib.og_forward_f32_via_f64.restype = ctypes.c_int
    return _lib


def _c_config(cfg: GenConfig) -> _OgConfig:
    c = _OgConfig()
    c.num_mels = 80
    c.upsample_initial_channel = cfg.upsample_initial_channel
    c.num_upsamples = len(cfg.upsample_rates)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        c.upsample_rates[i] = u
        c.upsample_kernel_sizes[i] = k
    c.num_kernels = len(cfg.resblock_kernel_sizes)
    nd = 3 if cfg.resblock == "1" else 2
    c.num_dilations = nd
    for j, (k, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
        assert len(ds) >= nd
        c.resblock_kernel_sizes[j] = k
        for m, d in enumerate(ds[:nd]):
            c.resblock_dilation_sizes[j][m] = d
    c.resblock_type = 1 if cfg.resblock == "1" else 2
    return c


def forward(cfg: GenConfig, state: Dict[str, np.ndarray], mel: np.ndarray, precision: str = "f32") -> np.ndarray:
    """Generator.forward on the CPU.  ``mel`` [B,80,T] -> [B,1,T*hop].

    precision: "f32" (fp32 storage + accumulation, like the reference), "f64" (all double;
    float64 in/out) or "f32_via_f64" (widen, compute in fp64, round once).
    """
    L = lib()
    ws = folded_weights(state, cfg)
    mel = np.asarray(mel)
    assert mel.ndim == 3 and mel.shape[1] == 80
    B, _, T = mel.shape
    dt = np.float64 if precision == "f64" else np.float32
    ws = [np.ascontiguousarray(w, dtype=dt) for w in ws]
    melc = np.ascontiguousarray(mel, dtype=dt)
    out = np.empty((B, 1, T * cfg.hop), dtype=dt)
    ptr_t = ctypes.POINTER(ctypes.c_double if dt == np.float64 else ctypes.c_float)
    arr = (ptr_t * len(ws))(*[w.ctypes.data_as(ptr_t) for w in ws])
    cc = _c_config(cfg)
    assert L.og_num_weight_arrays(ctypes.byref(cc)) == len(ws) == 2 * len(conv_specs(cfg))
    if precision == "f64":
        rc = L.og_forward_f64(ctypes.byref(cc), arr, melc.ctypes.data_as(ptr_t), B, T, out.ctypes.data_as(ptr_t))
    elif precision == "f32":
        rc = L.og_forward_f32(ctypes.byref(cc), arr, melc.ctypes.data_as(ptr_t), B, T, out.ctypes.data_as(ptr_t))
    elif precision == "f32_via_f64":
        sizes = (ctypes.c_int64 * len(ws))(*[w.size for w in ws])
        rc = L.og_forward_f32_via_f64(ctypes.byref(cc), arr, sizes, melc.ctypes.data_as(ptr_t), B, T,
                                      out.ctypes.data_as(ptr_t))
    else:
        raise ValueError(precision)
    if rc != 0:
        raise MemoryError("oracle allocation failed")
    return out


def conv1d(x, w, b, dilation: int, padding: int) -> np.ndarray:
    """Single same-length Conv1d in fp64 ([B,C,L] layout) — for op-level tests."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    B, cin, n = x.shape
    cout, _, k = w.shape
    y = np.empty((B, cout, n), dtype=np.float64)
    P = ctypes.POINTER(ctypes.c_double)
    L.og_op_conv1d_f64(x.ctypes.data_as(P), B, cin, n, w.ctypes.data_as(P), b.ctypes.data_as(P), cout, k,
                       dilation, padding, y.ctypes.data_as(P))
    return y


def conv_transpose1d(x, w, b, stride: int, padding: int) -> np.ndarray:
    """Single ConvTranspose1d in fp64 ([B,C,L] layout; weight [C_in,C_out,k])."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    B, cin, n = x.shape
    _, cout, k = w.shape
    lout = (n - 1) * stride - 2 * padding + k
    y = np.empty((B, cout, lout), dtype=np.float64)
    P = ctypes.POINTER(ctypes.c_double)
    L.og_op_conv_transpose1d_f64(x.ctypes.data_as(P), B, cin, n, w.ctypes.data_as(P), b.ctypes.data_as(P), cout, k,
                                 stride, padding, y.ctypes.data_as(P))
    return y


def to_int16(wav: np.ndarray, max_wav_value: float = 32768.0) -> np.ndarray:
    """HIFIapi.generate tail (hifiapi.py:50-51)."""
    L = lib()
    w = np.ascontiguousarray(wav, dtype=np.float32)
    out = np.empty(w.shape, dtype=np.int16)
    L.og_to_int16(w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_int64(w.size),
                  ctypes.c_float(max_wav_value), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)))
    return out
