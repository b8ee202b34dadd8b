"""CPU: libqdm.so loads, exports every symbol include/qdm.h declares, validates arguments before it
touches a device, and refuses to compute without a B200 (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "qdm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qdm_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(qdm):
    lib = qdm._lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libqdm.so does not export {s}"
    assert set(qdm._lib.SIGNATURES) == set(syms), set(qdm._lib.SIGNATURES) ^ set(syms)
    assert lib.qdm_version() >= 100


def test_argument_errors_map_to_reference_exception_types(qdm):
    lib = qdm._lib.load()
    # group must divide K -> QDM_ERR_INVALID -> ValueError (fake_quant.py:198,253 raise ValueError)
    rc = lib.qdm_quant_group(1, 0, 4, 100, 64, 4, 1, None, None, None, None, None, None, None, None)
    assert rc == qdm._lib.QDM_ERR_INVALID
    with pytest.raises(ValueError, match="group"):
        qdm._lib.check(rc)
    rc = lib.qdm_gemm_w4a16(16, 16, 16, 16, None, 16, 0, 128, 128, 100, 64, None)
    assert rc == qdm._lib.QDM_ERR_INVALID and "multiple of 64" in qdm._lib.last_error()
    rc = lib.qdm_pack_awq(16, 12, 64, 16, None)
    assert rc == qdm._lib.QDM_ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(qdm):
    lib = qdm._lib.load()
    assert lib.qdm_device_check(0) == qdm._lib.QDM_ERR_DEVICE
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        qdm.ops.colabsmax(torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        qdm._lib.check(lib.qdm_rowabsmax(16, 0, 4, 8, 16, None))


def test_new_entry_points_validate_arguments(qdm):
    """qdm_colstats / qdm_conv3x3_*: argument errors are reported before a device is touched."""
    lib, E = qdm._lib.load(), qdm._lib
    # no output requested / bad shape / bad mode
    assert lib.qdm_colstats(16, 0, 4, 8, 8, None, 0, None, None, 16, 1 << 20, None) == E.QDM_ERR_INVALID
    assert "no output" in E.last_error()
    assert lib.qdm_colstats(16, 0, 0, 8, 8, 16, 0, None, None, 16, 1 << 20, None) == E.QDM_ERR_INVALID
    assert lib.qdm_colstats(16, 0, 4, 8, 8, 16, 2, None, None, 16, 1 << 20, None) == E.QDM_ERR_INVALID
    assert lib.qdm_colstats_workspace_bytes(4096, 2432) == 2 * lib.qdm_colreduce_workspace_bytes(4096, 2432)
    # convolution geometry: C must be a multiple of 64; the direct form needs rows that tile 128
    assert lib.qdm_conv3x3_f16(16, 16, None, 16, 0, 1, 8, 8, 100, 64, None) == E.QDM_ERR_INVALID
    assert "multiple of 64" in E.last_error()
    assert lib.qdm_conv3x3_f16(16, 16, None, 16, 0, 0, 8, 8, 64, 64, None) == E.QDM_ERR_INVALID
    rc = lib.qdm_conv3x3_nhwc_f16(16, 16, None, 16, 0, 1, 5, 7, 64, 64, None)
    assert rc == E.QDM_ERR_UNSUPPORTED and "whole image rows" in E.last_error()
    with pytest.raises(RuntimeError, match="whole image rows"):
        E.check(rc)
    ok = lambda h, w: bool(lib.qdm_conv3x3_direct_ok(h, w))
    assert ok(64, 64) and ok(32, 32) and ok(16, 16) and ok(8, 8) and ok(128, 128) and ok(4, 4) and ok(16, 8)
    assert not ok(96, 96) and not ok(5, 7) and not ok(12, 16) and not ok(256, 256) and not ok(0, 8)
    # stride-2 form (the down-samplers): even input sizes whose OUTPUT grid tiles 128 rows; same channel rule
    assert lib.qdm_conv3x3s2_nhwc_f16(16, 16, None, 16, 0, 1, 15, 16, 64, 64, None) == E.QDM_ERR_INVALID
    assert "stride 2 needs even" in E.last_error()
    rc = lib.qdm_conv3x3s2_nhwc_f16(16, 16, None, 16, 0, 1, 16, 24, 64, 64, None)
    assert rc == E.QDM_ERR_UNSUPPORTED and "output grid 8 x 12" in E.last_error()
    assert lib.qdm_conv3x3s2_nhwc_f16(16, 16, None, 16, 0, 1, 16, 16, 100, 64, None) == E.QDM_ERR_INVALID
    assert lib.qdm_conv3x3s2_nhwc_w4a16(16, 16, 16, 16, None, 16, 0, 1, 15, 15, 64, 64, 64, None) == E.QDM_ERR_INVALID
    import importlib
    sys_ops = importlib.import_module("_oracle_ops")           # the CPU mirror used by the host-logic tests agrees with the library
    for h, w in ((64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2), (256, 256), (15, 16), (16, 24), (32, 16), (512, 512), (6, 6)):
        assert sys_ops.conv3x3_stride2_ok(h, w) == (h % 2 == 0 and w % 2 == 0 and ok(h // 2, w // 2)), (h, w)
    # packed-weight convolution: the group must be 64 * 2^j dividing 9 C
    assert lib.qdm_conv3x3_w4a16(16, 16, 16, 16, None, 16, 0, 1, 8, 8, 320, 64, 128, None) == E.QDM_ERR_INVALID
