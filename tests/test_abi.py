"""The C-ABI shared library: loads, exports every symbol include/hifigan_b200.h declares, and fails
loudly (no CPU fallback) when there is no sm_100 device.  No compute calls here."""
import ctypes
import os
import re

import pytest
import torch

from tts_king_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hifigan_b200.h")).read()
    return sorted(set(re.findall(r"HG_API\s+[\w\s\*]+?\b(hg_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("hg_plan_create", "hg_plan_upload_weight", "hg_plan_finalize", "hg_workspace_bytes", "hg_forward",
              "hg_plan_destroy", "hg_last_error", "hg_op_conv1d", "hg_op_conv_transpose1d", "hg_op_conv_post"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_native.lib_path())
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/hifigan_b200.h but not exported"
    assert _native.lib().hg_abi_version() == 1


def test_struct_layout_matches_header():
    # 3 + 8 + 8 + 1 + 8 + 8*4 + 1 int32 fields
    assert ctypes.sizeof(_native.HgConfig) == 4 * (3 + 8 + 8 + 1 + 8 + 32 + 1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_fails_loudly():
    L = _native.lib()
    assert L.hg_device_count() == 0
    cfg = _native.HgConfig()
    cfg.num_mels, cfg.upsample_initial_channel, cfg.num_upsamples, cfg.num_kernels, cfg.resblock_type = 80, 32, 1, 1, 1
    cfg.upsample_rates[0], cfg.upsample_kernel_sizes[0], cfg.resblock_kernel_sizes[0] = 2, 4, 3
    plan = ctypes.c_void_p()
    rc = L.hg_plan_create(ctypes.byref(cfg), 0, ctypes.byref(plan))
    assert rc == -2 and not plan.value
    assert b"no CPU fallback" in L.hg_last_error()
    with pytest.raises(_native.NativeError):
        _native.check(rc)


def test_bad_arguments_are_rejected_without_a_device():
    L = _native.lib()
    assert L.hg_plan_create(None, 0, None) == -1
    assert L.hg_plan_finalize(None) == -1
    n = ctypes.c_size_t()
    assert L.hg_workspace_bytes(None, 1, 1, 0, ctypes.byref(n)) == -1
