"""TEST INFRASTRUCTURE: a stand-in for the product's `ops` front end computed by the CPU oracle.

The product has no CPU path (ops.py rejects CPU tensors).  To exercise the HOST logic -- module swap, scale /
clip folding, packed checkpoints, the block-sharded search and its gather -- on the CPU-only box, the `-m "not gpu"`
tests patch `ops.<fn>` with the functions below for the duration of one test (`patched_ops()`); every function
is the oracle's restatement of the reference for the same arguments.  Nothing outside tests/ imports this file.
"""
import contextlib
import importlib

import numpy as np
import torch

import oracle.qdm_oracle as O

PKG = "quantization---diffusion-models_b200"


def _rows(x):
    return x.reshape(-1, x.shape[-1])


def colabsmax(x, out=None, running=False):
    m = O.hook_colabsmax(x)
    if running:
        out.copy_(torch.maximum(out, m))
        return out
    return m


def colabssum(x):
    return _rows(x).abs().float().sum(0)


def colstats(x, out_max=None, running=False, acc_maxsum=None, acc_abssum=None):
    m = O.hook_colabsmax(x)
    if acc_maxsum is not None:
        acc_maxsum += m.double()
    if acc_abssum is not None:
        acc_abssum += _rows(x).abs().float().sum(0).double()
    if out_max is None:
        return m if (acc_maxsum is None and acc_abssum is None) else None
    out_max.copy_(torch.maximum(out_max, m) if running else m)
    return out_max


def rowabsmax(x):
    return _rows(x).abs().amax(dim=-1)


def absmax(x):
    return x.abs().max()


def awq_wsum(w, group):
    g = w.reshape(-1, group)
    return (g.abs() / (g.abs().amax(dim=1, keepdim=True) + 1e-6)).view(w.shape).float().sum(0)


def sqdiff_sum(a, b, out=None):
    v = (a - b).float().pow(2).double().sum()
    if out is None:
        return v
    out.copy_(v)
    return out


def quant_group(w, group, n_bits=4, zero_point=True, no_clamp=False, pre_mul=None, clip_max=None, post_div=None,
                want_dq=True, want_codes=False, want_scales=True, out=None):
    k = w.shape[-1]
    group = k if group <= 0 else group
    if k % group:
        raise ValueError(f"group {group} must divide the last dim {k}")
    x = _rows(w)
    if pre_mul is not None:
        x = x * pre_mul.view(1, -1)
    if clip_max is not None:
        c = clip_max.reshape(x.shape[0], k // group, 1)
        x = torch.clamp(x.reshape(x.shape[0], k // group, group), -c, c).reshape(x.shape)
    if no_clamp:
        dq, codes, s = O.rtn_absmax_group(x, n_bits, group)
        dq, z = dq.to(w.dtype), None
        s = s.view(x.shape[0], -1)
    else:
        dq, s, z, codes = O.rtn_group(x, group, zero_point, n_bits)
    if post_div is not None:
        dq = dq / post_div.view(1, -1)
    dq = dq.reshape(w.shape)
    if out is not None:
        out.copy_(dq)
        dq = out
    codes = codes.reshape(w.shape).to(torch.uint8 if zero_point else torch.int8) if want_codes else None
    return (dq if want_dq else None), codes, (s if want_scales else None), (z if want_scales else None)


def quant_rowwise(x, n_bits=8, zero_point=False, no_clamp=True, want_dq=True, want_codes=False, want_scales=False):
    assert no_clamp and not zero_point
    dq, codes, s = O.rtn_rows(_rows(x), n_bits)
    return ((dq.reshape(x.shape) if want_dq else None),
            (codes.clamp(-128, 127).to(torch.int8).reshape(x.shape) if want_codes else None),
            (s.reshape(-1) if want_scales else None), None)


def quant_tensor(x, n_bits=8, want_dq=True, want_codes=False):
    dq, codes, s = O.rtn_tensor(x, n_bits)
    return (dq if want_dq else None), (codes.clamp(-128, 127).to(torch.int8) if want_codes else None), s


def actquant_token_i8(x, smooth=None):
    x2 = _rows(x)
    if smooth is not None:
        x2 = x2 / smooth.view(1, -1)
    _, codes, s = O.rtn_rows(x2, 8)
    return codes.clamp(-128, 127).to(torch.int8), s.reshape(-1).float()


def quant_pack_awq(w, group, want_dq=False):
    qw, qz, s, dq = O.awq_from_linear(w, group, 4)
    return torch.from_numpy(qw), torch.from_numpy(qz), s, (dq if want_dq else None)


def dequant_awq(qweight, qzeros, scales, group):
    return O.awq_dequant(qweight.numpy(), qzeros.numpy(), scales, group)


def pack_awq(codes_nk):
    return torch.from_numpy(O.awq_pack(codes_nk.t().contiguous().to(torch.int32).numpy()))


def unpack_awq(qweight):
    return torch.from_numpy(np.asarray(O.awq_unpack(qweight.numpy()))).to(torch.int8)


def _linear(x, w_nk, bias):
    y = torch.nn.functional.linear(x.float(), w_nk.float(), None if bias is None else bias.float())
    return y.to(x.dtype)


def gemm_f16(x, w, bias=None):
    return _linear(x, w, bias)


def gemm_f16_kn(x, w_kn, bias=None):
    return _linear(x, w_kn.t(), bias)


def awq_clip_search(w, x, group, n_bits=4, zero_point=True, n_grid=20, max_shrink=0.5):
    return O.awq_search_clip(w, x, group, zero_point, n_bits, n_grid, max_shrink, n_sample_token=1 << 30).unsqueeze(-1).reshape(w.shape[0], -1, 1)


def w4a16_repack(qweight, qzeros, scales, group):
    return torch.zeros(16, dtype=torch.uint8)        # the kernel-native copy only exists on the device


w4a16_repack_ts = w4a16_repack


def gemm_w4a16(x, qweight, qzeros, scales, group, bias=None, blob=None, blob_ts=None, out=None):
    y = _linear(x, dequant_awq(qweight, qzeros, scales, group).t(), bias)
    return y if out is None else out.copy_(y.reshape(out.shape))


def gemm_w8a8(xq, sx, wq, sw, bias=None, out_dtype=torch.float16):
    y = (xq.double() @ wq.double().t()) * sx.double()[:, None] * sw.double()[None, :]
    if bias is not None:
        y = y + bias.double()
    return y.to(out_dtype)


def geglu(x):
    h, gate = x.chunk(2, dim=-1)
    return h * torch.nn.functional.gelu(gate)


def conv3x3_weight_taps(w):
    n, c, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous()


def _taps_to_conv(w_tap, c):
    return w_tap.reshape(w_tap.shape[0], 3, 3, c).permute(0, 3, 1, 2)


def conv3x3_stride2_ok(h, w):
    wo, ho = w // 2, h // 2
    if h % 2 or w % 2 or wo <= 0 or wo > 128 or 128 % wo:
        return False
    rows = 128 // wo
    return ho % rows == 0 if rows <= ho else rows % ho == 0


def conv3x3_f16(x, w_tap, bias=None, padded=None, stride=1):
    return O.conv2d_fake(x, _taps_to_conv(w_tap, x.shape[1]), bias, stride, 1)


def conv3x3_w4a16(x, qweight, qzeros, scales, group, bias=None, padded=None, stride=1):
    return conv3x3_f16(x, dequant_awq(qweight, qzeros, scales, group).t(), bias, stride=stride)


NAMES = ("colabsmax", "colabssum", "colstats", "rowabsmax", "absmax", "awq_wsum", "sqdiff_sum", "quant_group",
         "quant_rowwise", "quant_tensor", "actquant_token_i8", "quant_pack_awq", "dequant_awq", "pack_awq", "unpack_awq",
         "awq_clip_search", "gemm_f16", "gemm_f16_kn", "gemm_w4a16", "w4a16_repack", "w4a16_repack_ts", "gemm_w8a8", "conv3x3_weight_taps", "conv3x3_f16", "conv3x3_w4a16", "conv3x3_stride2_ok",
         "geglu")


@contextlib.contextmanager
def patched_ops():
    """Swap the product's kernel front end for the oracle within a `with` block (CPU host-logic tests only)."""
    ops = importlib.import_module(PKG + ".ops")
    saved = {n: getattr(ops, n) for n in NAMES}
    try:
        for n in NAMES:
            setattr(ops, n, globals()[n])
        yield ops
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
