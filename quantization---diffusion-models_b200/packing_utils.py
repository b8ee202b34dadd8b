"""Mirror of the reference's utils/packing_utils.py (the int4 AWQ GEMM storage layout): same names and
argument meaning; the bit twiddling runs in libqdm kernels.  exllama repacking is out of scope."""
import torch

from . import ops

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]          # utils/packing_utils.py:4
AWQ_REVERSE_ORDER = [0, 4, 1, 5, 2, 6, 3, 7]  # utils/packing_utils.py:5


def _need_4bit(bits):
    if bits != 4:
        raise NotImplementedError(f"only the 4-bit AWQ layout exists on this path (got bits={bits})")


def unpack_awq(qweight: torch.Tensor, qzeros: torch.Tensor, bits: int):
    """packing_utils.py:8-26 followed by the order reversal of :29-43 is one kernel here, so this returns
    the codes still in PACKED nibble order exactly like the reference: column 8c+i = nibble i."""
    _need_4bit(bits)
    inv = torch.tensor(AWQ_ORDER, device=qweight.device)

    def packed_order(q):
        nat = ops.unpack_awq(q)                                   # natural column order
        return nat.view(nat.shape[0], -1, 8)[:, :, inv].reshape(nat.shape)

    return packed_order(qweight), (packed_order(qzeros) if qzeros is not None else None)


def reverse_awq_order(iweights: torch.Tensor, izeros: torch.Tensor, bits: int):
    """packing_utils.py:29-43 (pure index permutation)."""
    _need_4bit(bits)
    idx = torch.arange(iweights.shape[-1], dtype=torch.int64, device=iweights.device).view(-1, 8)[:, AWQ_REVERSE_ORDER].reshape(-1)
    if izeros is not None:
        izeros = izeros[:, idx]
    return iweights[:, idx], izeros


def dequantize_gemm(qweight, qzeros, scales, bits, group_size):
    """packing_utils.py:87-102: W_kn[K, N] = (q - z) * s in scales.dtype."""
    _need_4bit(bits)
    return ops.dequant_awq(qweight, qzeros, scales, group_size)
