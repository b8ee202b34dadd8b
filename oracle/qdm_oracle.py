"""CPU ORACLE for the quantized-linear hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch CPU tensors for the floating-point chains, numpy for the int4 byte
layout) of the reference algorithms the CUDA kernels must reproduce.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the
product package never does.

Pinned: tests/test_oracle_golden.py checks every function below against tests/golden/*.npz,
which tools/gen_golden.py produced by running the UNMODIFIED reference (imported from
/root/reference through tools/ref_shim.py) on seeded inputs; tests/test_oracle_vs_reference.py
repeats the comparison live on random inputs whenever /root/reference is present.

All `file:line` citations are relative to the reference root.  torch ops on fp16/bf16 CPU tensors
round to the tensor dtype after every op -- the same op-by-op rounding the reference's torch path
has on any device -- which is exactly what the kernels replay.
"""
import numpy as np
import torch

AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)          # utils/packing_utils.py:4, utils/quant_utils.py:10
AWQ_REVERSE_ORDER = (0, 4, 1, 5, 2, 6, 3, 7)  # utils/packing_utils.py:5


# ------------------------------------------------------------------ A1  quantize/quantizer.py:163-198
def rtn_group(w, group_size=128, zero_point=True, n_bits=4):
    """AwqQuantizer.pseudo_quantize_tensor. Returns (dq, scales, zeros|None, codes).
    `codes` (the integers before de-quantisation) are not returned by the reference; they are the
    clamp(...) term of quantizer.py:180 / :190, exposed so the kernels' int codes can be checked."""
    shape = w.shape
    g = w.reshape(-1, group_size) if group_size > 0 else w
    assert g.dim() == 2
    if zero_point:
        hi = g.amax(dim=1, keepdim=True)
        lo = g.amin(dim=1, keepdim=True)
        qmax, qmin = 2 ** n_bits - 1, 0
        scales = (hi - lo).clamp(min=1e-5) / qmax
        zeros = (-torch.round(lo / scales)).clamp_(qmin, qmax)
        codes = torch.clamp(torch.round(g / scales) + zeros, qmin, qmax)
        dq = (codes - zeros) * scales
        zeros = zeros.view(shape[0], -1)
    else:
        hi = g.abs().amax(dim=1, keepdim=True).clamp(min=1e-5)
        qmax, qmin = 2 ** (n_bits - 1) - 1, -(2 ** (n_bits - 1))
        scales = hi / qmax
        zeros = None
        codes = torch.clamp(torch.round(g / scales), qmin, qmax)
        dq = codes * scales
    return dq.reshape(shape), scales.view(shape[0], -1), zeros, codes.reshape(shape)


# ------------------------------------------------------------------ A2  quantize/fake_quant.py:21-84
def effective_group(k, group_size):
    """group-size fallback loop of fake_quant.py:34-37 (K=320 -> 64)."""
    while group_size > 0 and k % group_size != 0:
        group_size -= 32
    return group_size


def rtn_absmax_group(w, n_bits=8, group_size=0):
    """quantize_weight_absmax with codeBookQuantInd=False: symmetric, codes NOT clamped, result fp16.
    Works on a copy (the reference mutates its input view in place, fake_quant.py:72)."""
    shape = w.shape
    x = w.clone()
    if group_size > 0:
        group_size = effective_group(shape[-1], group_size)
        x = x.reshape(-1, group_size)
    assert x.dim() == 2
    s = x.abs().max(dim=-1, keepdim=True)[0]
    s.clamp_(min=1e-5).div_(2 ** (n_bits - 1) - 1)
    codes = x.div(s).round()
    dq = codes.mul(s)
    return dq.reshape(shape).to(torch.float16), codes.reshape(shape), s


# ------------------------------------------------------------------ A3/A5  fake_quant.py:86-93,109-118
def rtn_rows(x, n_bits=8):
    """quantize_weight_per_channel_absmax / quantize_activation_per_token_absmax: one scale per
    last-dim row, no clamp.  For a 4-D conv weight the rows are the kw taps (fake_quant.py:89)."""
    s = x.abs().max(dim=-1, keepdim=True)[0].clone()
    s.clamp_(min=1e-5).div_(2 ** (n_bits - 1) - 1)
    codes = x.div(s).round()
    return codes.mul(s).to(x.dtype), codes, s


# ------------------------------------------------------------------ A4  fake_quant.py:97-105,158-167
def rtn_tensor(x, n_bits=8):
    s = x.abs().max().clone()
    s.clamp_(min=1e-5).div_(2 ** (n_bits - 1) - 1)
    codes = x.div(s).round()
    return codes.mul(s).to(x.dtype), codes, s


# ------------------------------------------------------------------ A6  fake_quant.py:124-131
def rtn_nchw_channel(t, n_bits=8):
    s = torch.amax(t.abs(), dim=(2, 3), keepdim=True)
    s = s.clamp(min=1e-5) / (2 ** (n_bits - 1) - 1)
    return (t / s).round().mul(s).to(t.dtype)


def rtn_nchw_patch(t, group_size=128, n_bits=8):
    """quantize_activation_per_channel_group_absmax, fake_quant.py:134-153: one scale per (n, c, square spatial patch);
    the patch edge shrinks by 2 until it divides H and W (:140-141)."""
    n, c, h, w = t.shape
    while h % group_size != 0 or w % group_size != 0:
        group_size -= 2
    g = group_size
    p = t.reshape(n, c, h // g, g, w // g, g).permute(0, 1, 2, 4, 3, 5)            # [N, C, H/g, W/g, g, g] (the unfold view)
    s = torch.amax(p.abs(), dim=(4, 5), keepdim=True).clamp(min=1e-5) / (2 ** (n_bits - 1) - 1)
    q = p.div(s).round().mul(s)
    return q.permute(0, 1, 2, 4, 3, 5).reshape(n, c, h, w).to(t.dtype)


# ------------------------------------------------------------------ A15 int4 AWQ GEMM layout
def awq_pack(codes_kn):
    """codes [K, N] (0..15) -> int32 [K, N/8]; nibble i of a word = column 8c + AWQ_ORDER[i]
    (quant_utils.py:14-39 `pack(apply_order(., AWQ_PACK_ORDER))`; inverse of packing_utils.py:8-43)."""
    c = np.asarray(codes_kn).astype(np.uint32) & 0xF
    k, n = c.shape
    c = c.reshape(k, n // 8, 8)[:, :, list(AWQ_ORDER)]
    word = np.zeros((k, n // 8), dtype=np.uint32)
    for i in range(8):
        word |= c[:, :, i] << np.uint32(4 * i)
    return word.view(np.int32)


def awq_unpack(qweight):
    """int32 [K, N/8] -> codes uint8 [K, N] in natural column order (packing_utils.py:8-43)."""
    q = np.asarray(qweight).view(np.uint32)
    k, nw = q.shape
    out = np.zeros((k, nw, 8), dtype=np.uint8)
    for i in range(8):
        out[:, :, AWQ_ORDER[i]] = (q >> np.uint32(4 * i)) & 0xF
    return out.reshape(k, nw * 8)


def awq_dequant(qweight, qzeros, scales, group_size):
    """dequantize_gemm, packing_utils.py:87-102: W_kn = (q - z) * s, scales.dtype."""
    q = torch.from_numpy(awq_unpack(qweight).astype(np.int8))
    z = torch.from_numpy(awq_unpack(qzeros).astype(np.int8))
    s = scales.repeat_interleave(group_size, dim=0)
    z = z.repeat_interleave(group_size, dim=0)
    return (q - z) * s


def awq_from_linear(w, group_size=128, n_bits=4):
    """pseudo_quantize_tensor + transpose + pack, quantizer.py:540-569: returns
    (qweight [K,N/8], qzeros [K/g,N/8], scales [K/g,N], dq [N,K])."""
    dq, scales, zeros, codes = rtn_group(w, group_size, True, n_bits)
    qweight = awq_pack(codes.t().contiguous().to(torch.int32).numpy())
    qzeros = awq_pack(zeros.t().contiguous().to(torch.int32).numpy())
    return qweight, qzeros, scales.t().contiguous(), dq


# ------------------------------------------------------------------ A9  quantizer.py:627-659
def awq_w_mean(weights, group_size):
    w = torch.cat(list(weights), dim=0)
    shape = w.shape
    g = w.view(-1, group_size)
    w_scale = g.abs() / (g.abs().amax(dim=1, keepdim=True) + 1e-6)
    return w_scale.view(shape).mean(0)


def awq_x_mean(inp):
    flat = inp.abs().view(-1, inp.shape[-1])
    x_sum = flat.to(torch.float32).sum(dim=0)
    return (x_sum / flat.size(0)).to(inp.dtype)


# ------------------------------------------------------------------ A10 quantizer.py:678-783
def awq_ratio_scales(x_mean, w_mean, ratio, duo_scaling=True):
    if duo_scaling:
        s = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
    else:
        s = x_mean.pow(ratio).clamp(min=1e-4).view(-1)
    s = s / (s.max() * s.min()).sqrt()
    s[torch.isinf(s)] = 1
    s[torch.isnan(s)] = 1
    return s


def mse_loss(ref_out, out):
    """_compute_loss, quantizer.py:754-783 (single chunk)."""
    return (ref_out.view(-1) - out.view(-1)).float().pow(2).sum().item() / ref_out.numel()


def awq_search_scale(x, weights, forward, group_size=128, zero_point=True, n_bits=4, duo_scaling=True, n_grid=20):
    """_search_best_scale/_compute_best_scale for a group of Linear weights sharing input x.
    `forward(list_of_weights)` evaluates the inspected module on x.  Returns (best_scales, best_ratio, losses)."""
    w_mean = awq_w_mean(weights, group_size)
    x_mean = awq_x_mean(x)
    ref_out = forward(list(weights))
    best, best_ratio, best_s, hist = float("inf"), -1, None, []
    for i in range(n_grid):
        ratio = i / n_grid
        s = awq_ratio_scales(x_mean.view(-1), w_mean.view(-1), ratio, duo_scaling)
        sv = s.view(1, -1)
        qws = [rtn_group(w * sv, group_size, zero_point, n_bits)[0] / sv for w in weights]
        loss = mse_loss(ref_out, forward(qws))
        hist.append(loss)
        if loss < best:
            best, best_ratio, best_s = loss, ratio, s.clone()
    return best_s, best_ratio, hist


# ------------------------------------------------------------------ A11 quantizer.py:805-863
def awq_search_clip(w, x, group_size=128, zero_point=True, n_bits=4, n_grid=20, max_shrink=0.5, n_sample_token=512,
                    return_levels=False):
    """return_levels: also ([levels, co, G] candidate max values, [levels, co, G] errors) as the reference computes them
    (tests use them to judge a near-tie: is the level another implementation picked as good as the best one?)."""
    co = w.shape[0]
    gs = group_size if group_size > 0 else w.shape[1]
    x = x.view(-1, x.shape[-1]).reshape(1, -1, w.shape[1] // gs, gs)
    x = x[:, :: max(1, x.shape[1] // n_sample_token)]
    w4 = w.reshape(co, 1, -1, gs)
    bs = 256 if co % 256 == 0 else 64
    assert co % bs == 0
    outs, lv_max, lv_err = [], [], []
    for b in range(co // bs):
        wb = w4[b * bs:(b + 1) * bs]
        org_max = wb.abs().amax(dim=-1, keepdim=True)
        best_max = org_max.clone()
        min_errs = torch.ones_like(org_max) * 1e9
        org_out = (x * wb).sum(dim=-1)
        ms, es = [], []
        for i_s in range(int(max_shrink * n_grid)):
            max_val = org_max * (1 - i_s / n_grid)
            cur = torch.clamp(wb, -max_val, max_val)
            q = rtn_group(cur, gs, zero_point, n_bits)[0]
            err = ((x * q).sum(dim=-1) - org_out).pow(2).mean(dim=1).view(min_errs.shape)
            better = err < min_errs
            min_errs[better] = err[better]
            best_max[better] = max_val[better]
            ms.append(max_val.squeeze(1).squeeze(-1)), es.append(err.squeeze(1).squeeze(-1))
        outs.append(best_max)
        lv_max.append(torch.stack(ms)), lv_err.append(torch.stack(es))
    best = torch.cat(outs, dim=0).squeeze(1)
    if return_levels:
        return best, torch.cat(lv_max, dim=1), torch.cat(lv_err, dim=1)
    return best


def apply_clip(w, max_val):
    """quantize/scale.py:25-34."""
    shape = w.shape
    return torch.clamp(w.reshape(*max_val.shape[:2], -1), -max_val, max_val).reshape(shape)


# ------------------------------------------------------------------ A13 utils/calib_data.py:112-121
def hook_colabsmax(x):
    return x.reshape([-1, x.shape[-1]]).abs().amax(dim=0)


def mean_of_calls(per_call_max):
    """mean over calls of the per-call maxima, models/StableDiffusion1_x.py:104-112."""
    return torch.stack(list(per_call_max)).mean(dim=0)


# ------------------------------------------------------------------ A14 quantizer_SQ.py:396-431
def smooth_scales(act_scales, weights, alpha=0.5):
    w_s = torch.cat([w.abs().max(dim=0, keepdim=True)[0] for w in weights], dim=0)
    w_s = w_s.max(dim=0)[0].clamp(min=1e-5)
    return (act_scales.pow(alpha) / w_s.pow(1 - alpha)).clamp(min=1e-5)


def smooth_fold(ln_weight, ln_bias, weights, s):
    return ln_weight / s, (ln_bias / s if ln_bias is not None else None), [w * s.view(1, -1) for w in weights]


# ------------------------------------------------------------------ A7  fake_quant.py:215-225
def linear_fake(x, w_fake, bias=None):
    """WxAxLinear.forward with quantize_act=False: F.linear on fake-quant weights (fp32 math on CPU)."""
    y = torch.nn.functional.linear(x.float(), w_fake.float(), None if bias is None else bias.float())
    return y.to(x.dtype)


def linear_w8a8_fake(x, w, bias=None):
    """W8A8 in the reference's fake-quant formulation: per-token A8 (fake_quant.py:109-118) x per-channel
    W8 (fake_quant.py:86-93), then F.linear."""
    xq = rtn_rows(x.reshape(-1, x.shape[-1]), 8)[0]
    wq = rtn_rows(w, 8)[0]
    return linear_fake(xq, wq, bias).reshape(*x.shape[:-1], w.shape[0])


# ------------------------------------------------------------------ A8  fake_quant.py:337-341
def conv2d_fake(x, w_fake, bias=None, stride=1, padding=0):
    """WxAxConv2d.forward with quantize_act=False: F.conv2d on fake-quant weights (fp32 math on CPU)."""
    y = torch.nn.functional.conv2d(x.float(), w_fake.float(), None if bias is None else bias.float(), stride, padding)
    return y.to(x.dtype)
