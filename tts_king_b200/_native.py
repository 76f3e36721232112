"""ctypes binding of ``libhifigan_b200.so`` (C ABI: ``include/hifigan_b200.h``).

There is no fallback: if the shared library is missing this module raises at first use, and every
compute call fails loudly when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

HG_MAX_UPS, HG_MAX_KERNELS, HG_MAX_DIL = 8, 8, 4
PREC_BF16, PREC_FP32, PREC_FP32_FFMA = 0, 1, 2
OUT_F32, OUT_I16 = 0, 1
PRECISIONS = {"bf16": PREC_BF16, "fp32": PREC_FP32, "fp32_ffma": PREC_FP32_FFMA}

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libhifigan_b200.so")
_lib: Optional[ctypes.CDLL] = None


class HgConfig(ctypes.Structure):
    _fields_ = [
        ("num_mels", ctypes.c_int32),
        ("upsample_initial_channel", ctypes.c_int32),
        ("num_upsamples", ctypes.c_int32),
        ("upsample_rates", ctypes.c_int32 * HG_MAX_UPS),
        ("upsample_kernel_sizes", ctypes.c_int32 * HG_MAX_UPS),
        ("num_kernels", ctypes.c_int32),
        ("resblock_kernel_sizes", ctypes.c_int32 * HG_MAX_KERNELS),
        ("resblock_dilation_sizes", (ctypes.c_int32 * HG_MAX_DIL) * HG_MAX_KERNELS),
        ("resblock_type", ctypes.c_int32),
    ]


class HgLayerInfo(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64)] + [(n, ctypes.c_int32) for n in (
        "kind", "c_in", "c_out", "k", "dilation", "stride", "tensor_core", "n_tile", "k_chunk", "m_subtiles",
        "stages", "smem_bytes", "weights_resident", "slab_buffers", "kernel_path")]


class HgFoldInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "fusable", "f", "r_out", "delta", "fdiv", "blk_off", "nb_slab", "slab_phase_bytes", "xt_phase_bytes", "t_bufs",
        "stages", "weights_resident", "smem_bytes", "n_ops1", "n_ops2")] + [
        ("ops1", (ctypes.c_int32 * 6) * 24), ("ops2", (ctypes.c_int32 * 6) * 24)]


KERNEL_PATHS = {0: "cuda-core", 1: "tcgen05", 2: "tcgen05 cta_group::2", 3: "tcgen05 fused pair", 4: "cuda-core narrow",
                5: "conv_post", 6: "repack", 7: "tcgen05 fused ResBlock"}


class HgStackLayer(ctypes.Structure):
    _fields_ = [("c_in", ctypes.c_int32), ("c_out", ctypes.c_int32), ("k", ctypes.c_int32), ("dilation", ctypes.c_int32),
                ("act", ctypes.c_int32), ("slope", ctypes.c_float)]


ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2  # HG_ACT_*
MAX_RAGGED_ITEMS = 64  # kMaxRaggedItems, csrc/common.cuh (hg_forward_ragged's bound on B)


class NativeError(RuntimeError):
    """Raised for every non-zero status from the C ABI (the reference raises RuntimeError too)."""


def lib_path() -> str:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeError(
            f"{_LIB_PATH} is missing — build it with `python -m tts_king_b200.build` "
            "(or __graft_entry__.build()). tts_king_b200 has no CPU/PyTorch fallback."
        )
    L = ctypes.CDLL(_LIB_PATH)
    vp, i, i64, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
    L.hg_abi_version.restype = i
    L.hg_last_error.restype = ctypes.c_char_p
    L.hg_device_count.restype = i
    L.hg_plan_create.argtypes = [ctypes.POINTER(HgConfig), i, ctypes.POINTER(vp)]
    L.hg_plan_upload_weight.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i, vp, i64]
    L.hg_plan_finalize.argtypes = [vp]
    L.hg_workspace_bytes.argtypes = [vp, i, i, i, ctypes.POINTER(sz)]
    L.hg_forward_launches.argtypes = [vp, i, i, i, ctypes.POINTER(i)]
    L.hg_forward.argtypes = [vp, vp, i64, i64, i64, i, i, vp, i, f, i, vp, sz, vp]
    L.hg_forward_window.argtypes = [vp, vp, i64, i64, i64, i, i, vp, i64, i64, i64, i, f, i, vp, sz, vp]
    L.hg_enable_peer_access.argtypes = [i, i]
    L.hg_ipc_export.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(i64)]
    L.hg_ipc_import.argtypes = [i, ctypes.c_char_p, i64, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.hg_ipc_close.argtypes = [i, vp]
    L.hg_halo_frames.argtypes = [vp, ctypes.POINTER(i)]
    L.hg_forward_ragged.argtypes = [vp, vp, i64, i64, i64, i, i, ctypes.POINTER(ctypes.c_int32), vp, i, f, i, vp, sz, vp]
    L.hg_stack_create.argtypes = [ctypes.POINTER(HgStackLayer), i, i, ctypes.POINTER(vp)]
    L.hg_stack_workspace_bytes.argtypes = [vp, i, i, i, ctypes.POINTER(sz)]
    L.hg_stack_forward.argtypes = [vp, vp, i64, i64, i64, i, i, vp, vp, i, vp, sz, vp]
    L.hg_plan_destroy.argtypes = [vp]
    L.hg_op_conv1d.argtypes = [i, i, vp, i, i, i, vp, vp, i, i, i, f, vp, vp, vp]
    L.hg_op_conv_transpose1d.argtypes = [i, i, vp, i, i, i, vp, vp, i, i, i, f, vp, vp]
    L.hg_op_conv_post.argtypes = [i, vp, i, i, i, vp, vp, vp, vp]
    L.hg_selftest_tcgen05.argtypes = [i, ctypes.c_char_p, sz]
    L.hg_op_conv_pair.argtypes = [i, vp, i, i, i, i, i, vp, vp, vp, vp, f, vp, vp, vp]
    pp = ctypes.POINTER(vp)
    L.hg_op_resblock1.argtypes = [i, vp, i, i, i, i, i, ctypes.POINTER(ctypes.c_int32), pp, pp, pp, pp, f, vp, vp,
                                  ctypes.POINTER(ctypes.c_int32)]
    L.hg_op_resblock1.restype = i
    L.hg_fold_info.argtypes = [i, i, i, i, ctypes.POINTER(HgFoldInfo)]
    L.hg_fold_info.restype = i
    p32 = ctypes.POINTER(ctypes.c_int32)
    L.hg_fold_ring_query.argtypes = [i, i, i, i, p32, p32, p32, p32]
    L.hg_fold_ring_query.restype = i
    L.hg_layer_count.argtypes = [vp, ctypes.POINTER(i)]
    L.hg_layer_info.argtypes = [vp, i, i, ctypes.POINTER(HgLayerInfo)]
    L.hg_profile_launch_info.argtypes = [vp, i, ctypes.POINTER(HgLayerInfo)]
    L.hg_profile_forward.argtypes = [vp, vp, i64, i64, i64, i, i, vp, i, f, i, vp, sz, vp, ctypes.POINTER(i),
                                     ctypes.POINTER(f), i, ctypes.POINTER(i)]
    for name in ("hg_plan_create", "hg_plan_upload_weight", "hg_plan_finalize", "hg_workspace_bytes",
                 "hg_forward_launches", "hg_forward", "hg_halo_frames", "hg_forward_ragged", "hg_forward_window", "hg_enable_peer_access", "hg_ipc_export", "hg_ipc_import", "hg_ipc_close", "hg_stack_create",
                 "hg_stack_workspace_bytes", "hg_stack_forward", "hg_plan_destroy", "hg_op_conv1d",
                 "hg_op_conv_transpose1d", "hg_op_conv_post", "hg_op_conv_pair", "hg_selftest_tcgen05", "hg_layer_count",
                 "hg_layer_info", "hg_profile_launch_info",
                 "hg_profile_forward"):
        getattr(L, name).restype = i
    if L.hg_abi_version() != 1:
        raise NativeError("libhifigan_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().hg_last_error().decode("utf-8", "replace")
        raise NativeError(f"libhifigan_b200: {msg} (status {rc})")


def selftest(device: int = 0) -> tuple[int, str]:
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().hg_selftest_tcgen05(device, buf, len(buf))
    return n, buf.value.decode()
