"""Mirror of the reference's utils/quant_utils.py pack/unpack helpers for the AWQ ("column") direction."""
from typing import List

import torch

from . import ops

Q_BITS = 4
STORAGE_BITS = 32
PACK_NUM = STORAGE_BITS // Q_BITS
ORDINAL_PACK_ORDER = [0, 1, 2, 3, 4, 5, 6, 7]
AWQ_PACK_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]
REVERSE_AWQ_PACK_ORDER = [0, 4, 1, 5, 2, 6, 3, 7]


def apply_order(imatrix: torch.Tensor, direction: str = "column", order: List[int] = ORDINAL_PACK_ORDER):
    """quant_utils.py:122-144 (index permutation only)."""
    if direction == "column":
        return imatrix.view(-1, PACK_NUM)[:, order].view(imatrix.shape)
    if direction == "row":
        return imatrix.view(PACK_NUM, -1)[order, :].view(imatrix.shape)
    raise ValueError(f"Invalid direction: {direction}")


def pack_awq(imatrix_kn: torch.Tensor):
    """== pack(apply_order(imatrix, "column", AWQ_PACK_ORDER), "column") of quant_utils.py:14-39,122-144:
    codes [K, N] -> int32 [K, N/8] in one kernel (the kernel reads the [N, K] orientation)."""
    return ops.pack_awq(imatrix_kn.to(torch.int8).t().contiguous())


def unpack_awq(qmatrix: torch.Tensor):
    """== apply_order(unpack(qmatrix, "column"), "column", REVERSE_AWQ_PACK_ORDER): natural-order codes [K, N]."""
    return ops.unpack_awq(qmatrix)


def dequantize(imatrix, scales, zeros, group_size):
    """quant_utils.py:97-119 on already-unpacked natural-order codes (plain tensor algebra, layout check helper)."""
    z = (zeros.to(torch.int8) & 0x0F).repeat_interleave(group_size, dim=0)
    q = imatrix.to(torch.int8) & 0x0F
    return ((q - z) * scales.repeat_interleave(group_size, dim=0)).to(torch.float16)
