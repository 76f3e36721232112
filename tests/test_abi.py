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


def test_struct_layout_matches_header(tmp_path):
    # 3 + 8 + 8 + 1 + 8 + 8*4 + 1 int32 fields
    assert ctypes.sizeof(_native.HgConfig) == 4 * (3 + 8 + 8 + 1 + 8 + 32 + 1)
    # the header itself, through a C compiler: it must be plain C (no C++-isms) and the ctypes mirrors of
    # its structs must have the same size and field offsets
    import shutil
    import subprocess

    cc = shutil.which("gcc", path="/usr/bin:/bin") or shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "hifigan_b200.h"\n'
        "int main(void) {\n"
        '  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(HgConfig), sizeof(HgLayerInfo), sizeof(HgStackLayer),\n'
        "         offsetof(HgConfig, resblock_dilation_sizes), offsetof(HgLayerInfo, kernel_path),\n"
        "         offsetof(HgLayerInfo, n_tile), offsetof(HgStackLayer, slope), sizeof(HgFoldInfo),\n"
        "         offsetof(HgFoldInfo, ops1), offsetof(HgFoldInfo, ops2));\n"
        "  return (HG_PATH_REPACK == 6 && HG_ACT_TANH == 2 && HG_OUT_I16 != HG_OUT_F32) ? 0 : 1;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    want = [ctypes.sizeof(_native.HgConfig), ctypes.sizeof(_native.HgLayerInfo), ctypes.sizeof(_native.HgStackLayer),
            _native.HgConfig.resblock_dilation_sizes.offset, _native.HgLayerInfo.kernel_path.offset,
            _native.HgLayerInfo.n_tile.offset, _native.HgStackLayer.slope.offset, ctypes.sizeof(_native.HgFoldInfo),
            _native.HgFoldInfo.ops1.offset, _native.HgFoldInfo.ops2.offset]
    assert [int(v) for v in out] == want


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
