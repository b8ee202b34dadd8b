"""ctypes binding of libqdm.so (C ABI declared in include/qdm.h).

This is the only place the shared library is touched.  There is no fallback: if the
library is missing the import raises, and every entry point refuses to run on anything
that is not a cc-10.0 (B200) device.
"""
import ctypes
import os
from ctypes import c_char_p, c_int, c_int64, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QDM_LIB", os.path.join(_HERE, "libqdm.so"))   # QDM_LIB: debug builds (timeline tracing)

QDM_OK = 0
QDM_ERR_INVALID = -1
QDM_ERR_CUDA = -2
QDM_ERR_UNSUPPORTED = -3
QDM_ERR_DEVICE = -4

QDM_F16, QDM_BF16, QDM_F32 = 0, 1, 2
QDM_Q_ZERO_POINT, QDM_Q_NO_CLAMP = 1, 2

_P = c_void_p
_I = c_int
_L = c_int64
_Z = c_size_t
_U = c_uint

# name -> (restype, argtypes); mirrors include/qdm.h line by line
SIGNATURES = {
    "qdm_version": (c_int, []),
    "qdm_last_error": (c_char_p, []),
    "qdm_device_check": (c_int, [_I]),
    "qdm_launch_count": (c_int64, [_I]),
    "qdm_colreduce_workspace_bytes": (c_size_t, [_L, _L]),
    "qdm_colabsmax": (c_int, [_P, _I, _L, _L, _L, _P, _I, _P, _Z, _P]),
    "qdm_colabssum": (c_int, [_P, _I, _L, _L, _L, _P, _P, _Z, _P]),
    "qdm_colstats_workspace_bytes": (c_size_t, [_L, _L]),
    "qdm_colstats": (c_int, [_P, _I, _L, _L, _L, _P, _I, _P, _P, _P, _Z, _P]),
    "qdm_rowabsmax": (c_int, [_P, _I, _L, _L, _P, _P]),
    "qdm_absmax_workspace_bytes": (c_size_t, [_L]),
    "qdm_absmax": (c_int, [_P, _I, _L, _P, _P, _Z, _P]),
    "qdm_awq_wsum": (c_int, [_P, _I, _L, _L, _I, _P, _P, _Z, _P]),
    "qdm_sqdiff_workspace_bytes": (c_size_t, [_L]),
    "qdm_sqdiff_sum": (c_int, [_P, _P, _I, _L, _P, _P, _Z, _P]),
    "qdm_awq_clip_workspace_bytes": (c_size_t, [_L, _I]),
    "qdm_awq_clip_search": (c_int, [_P, _I, _L, _L, _I, _I, _U, _P, _L, _L, _I, ctypes.c_float, _P, _P, _Z, _P]),
    "qdm_quant_group": (c_int, [_P, _I, _L, _L, _I, _I, _U, _P, _P, _P, _P, _P, _P, _P, _P]),
    "qdm_quant_rowwise": (c_int, [_P, _I, _L, _L, _I, _U, _P, _P, _P, _P, _P]),
    "qdm_quant_tensor_workspace_bytes": (c_size_t, [_L]),
    "qdm_quant_tensor": (c_int, [_P, _I, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "qdm_pack_awq": (c_int, [_P, _L, _L, _P, _P]),
    "qdm_unpack_awq": (c_int, [_P, _L, _L, _P, _P]),
    "qdm_quant_pack_awq": (c_int, [_P, _I, _L, _L, _I, _P, _P, _P, _P, _P]),
    "qdm_dequant_awq": (c_int, [_P, _P, _P, _I, _L, _L, _I, _P, _P]),
    "qdm_actquant_token_i8": (c_int, [_P, _I, _L, _L, _P, _P, _P, _P]),
    "qdm_geglu": (c_int, [_P, _I, _L, _L, _P, _P]),
    "qdm_set_gemm_mode": (c_int, [_I]),
    "qdm_gemm_f16": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _P]),
    "qdm_gemm_f16_kn": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _P]),
    "qdm_gemm_w4a16": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _I, _P]),
    "qdm_gemm_w8a8": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _P]),
    "qdm_w4a16_repack_bytes": (c_size_t, [_L, _L]),
    "qdm_w4a16_repack": (c_int, [_P, _P, _P, _L, _L, _I, _P, _Z, _P]),
    "qdm_gemm_w4a16_rp": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _I, _P]),
    "qdm_gemm_last_variant": (c_int, [_P]),
    "qdm_w4a16_repack_ts_bytes": (c_size_t, [_L, _L]),
    "qdm_w4a16_repack_ts": (c_int, [_P, _P, _P, _I, _L, _L, _I, _P, _Z, _P]),
    "qdm_gemm_w4a16_plan": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _I, _P]),
    "qdm_conv3x3_f16": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _P]),
    "qdm_conv3x3_w4a16": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _I, _P]),
    "qdm_conv3x3_direct_ok": (c_int, [_L, _L]),
    "qdm_conv3x3_nhwc_f16": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _P]),
    "qdm_conv3x3_nhwc_w4a16": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _I, _P]),
    "qdm_conv3x3s2_nhwc_f16": (c_int, [_P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _P]),
    "qdm_conv3x3s2_nhwc_w4a16": (c_int, [_P, _P, _P, _P, _P, _P, _I, _L, _L, _L, _L, _L, _I, _P]),
    "qdm_set_w4_disable": (c_int, [_I]),
    "qdm_selftest_fastdiv": (c_int, [_I, _P]),
    "qdm_gemm_workspace_bytes": (c_size_t, []),
    "qdm_gemm_set_workspace": (c_int, [_P, _Z, _P]),
}

_lib = None


def load():
    """Load libqdm.so once; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library is stale
        fn.restype = res
        fn.argtypes = args
    mode = os.environ.get("QDM_GEMM_MODE")   # bring-up / A-B timing switch, see qdm_set_gemm_mode
    if mode:
        lib.qdm_set_gemm_mode(int(mode))
    _lib = lib
    return lib


def last_error():
    msg = load().qdm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Map return codes to the reference's exception types (ValueError for bad arguments:
    quantize/fake_quant.py:198,253; RuntimeError otherwise)."""
    if rc == QDM_OK:
        return
    msg = last_error()
    if rc == QDM_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"libqdm error {rc}: {msg}")
