"""B200-native quantized-linear hot path of maani3/Quantization---Diffusion-Models.

Python host code mirrors the reference's operator API (quantize/fake_quant.py, quantize/quantizer.py,
quantize/quantizer_SQ.py, quantize/scale.py, utils/packing_utils.py); all arithmetic runs in
hand-written sm_100a kernels behind the C ABI of include/qdm.h (libqdm.so, loaded with ctypes).
"""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
