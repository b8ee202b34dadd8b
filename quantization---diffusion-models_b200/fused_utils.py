"""Same-input packed projections as ONE W4A16 (or W8A8) launch.

The AWQ GEMM layout (utils/packing_utils.py:37-102) keeps the output channels on the last axis of all three tensors
(qweight [K, N/8], qzeros [K/g, N/8], scales [K/g, N]) and packs 8 ADJACENT channels into a word, so the packed
tensors of Linears that read the same input concatenate along dim 1 without touching a code: the reference's
`fuse_qkv` (utils/fused_utils.py:45-143, GEMM branch :87-96) and `fuse_linears` (:145-163) do exactly that for
`WQLinear_GEMM`.  Here the same two entry points build the fused module, and `fuse_projections` applies them to the
denoisers this repo runs:

  * self-attention  to_q / to_k / to_v                (UNet `attn1`, MMDiT `attn` and its `add_*_proj`)   -> one launch
  * cross-attention to_k / to_v of EVERY block        (all read the prompt embedding)                      -> one launch per step
  * `time_emb_proj` of every ResNet block             (all read silu(temb))                                -> one launch per step
  * AdaLN modulation Linears of every MMDiT block     (all read silu(emb))                                 -> one launch per step
  * GEGLU between ff.net.0.proj and ff.net.2          (gelu + mul and two temporaries)                     -> ops.geglu

A denoise step of the SD1.5 UNet goes from 184 to 100 Linear launches with the same FLOPs; the activations that were
read three times are read once and the latency-bound M = 16 / M = 1232 launches (54 of the 184) become two.
The member modules keep their own buffers (state dict, save_quantized and the per-layer parity tests are unchanged);
the fused copy costs another 0.5 B / weight of the fused members.
"""
import torch
import torch.nn.functional as F

from .linear import W8A8Linear, WQLinear_GEMM


def _cat_bias(linears, dev, dtype):
    if not any(l.bias is not None for l in linears):
        return None
    return torch.cat([l.bias if l.bias is not None else torch.zeros(l.out_features, device=dev, dtype=dtype) for l in linears])


def _fuse_w8a8(linears):
    """W8A8Linear members (int8 codes [N, K], one fp32 scale per output row): rows concatenate; the shared input is quantised
    per token ONCE for all members (quantize/fake_quant.py:109-118 runs once per member in the reference)."""
    first = linears[0]
    for l in linears:
        if not isinstance(l, W8A8Linear) or (l.in_features, l.out_dtype) != (first.in_features, first.out_dtype):
            raise ValueError("fuse_linears: W8A8 members differ in in_features / dtype")
        if (l.smooth is None) != (first.smooth is None) or (l.smooth is not None and not torch.equal(l.smooth, first.smooth)):
            raise ValueError("fuse_linears: W8A8 members must share the SmoothQuant activation divisor")
    dev = first.qweight.device
    fused = W8A8Linear(first.in_features, sum(l.out_features for l in linears), False, dev, first.out_dtype)
    fused.qweight = torch.cat([l.qweight for l in linears], dim=0).contiguous()
    fused.w_scales = torch.cat([l.w_scales for l in linears]).contiguous()
    fused.smooth = first.smooth
    fused.bias = _cat_bias(linears, dev, first.out_dtype)
    fused.split_sizes = [l.out_features for l in linears]
    return fused


def fuse_linears(linears, device=None, dim=1, operation=torch.cat):
    """utils/fused_utils.py:145-163 -- `linears` read the same input; returns one WQLinear_GEMM with
    out_features = sum.  Unlike the reference the members are left intact and a bias is carried (zeros for a member
    that has none), so that the fused module is a drop-in for the members' concatenated outputs.
    W8A8Linear members (this repo's kernel (d) module) fuse the same way along their output rows."""
    first = linears[0]
    if isinstance(first, W8A8Linear):
        return _fuse_w8a8(linears)
    for l in linears:
        if not isinstance(l, WQLinear_GEMM):
            raise TypeError(f"fuse_linears: {type(l).__name__} is not a WQLinear_GEMM")
        if (l.in_features, l.group_size, l.w_bit, l.scales.dtype) != (first.in_features, first.group_size, first.w_bit, first.scales.dtype):
            raise ValueError("fuse_linears: members differ in in_features / group_size / w_bit / dtype")
    if dim != 1:
        raise ValueError("fuse_linears: GEMM-format tensors concatenate along dim 1 (utils/fused_utils.py:87-96)")
    dev = first.qweight.device if device is None else device
    has_bias = any(l.bias is not None for l in linears)
    fused = WQLinear_GEMM(first.w_bit, first.group_size, first.in_features, sum(l.out_features for l in linears), has_bias, dev,
                          first.scales.dtype)
    fused.qweight = operation([l.qweight for l in linears], dim=1).contiguous()
    fused.qzeros = operation([l.qzeros for l in linears], dim=1).contiguous()
    fused.scales = operation([l.scales for l in linears], dim=1).contiguous()
    if has_bias:
        fused.bias = torch.cat([l.bias if l.bias is not None else torch.zeros(l.out_features, device=dev, dtype=first.scales.dtype)
                                for l in linears])
    fused.split_sizes = [l.out_features for l in linears]
    return fused


def fuse_qkv(module, q_proj, k_proj, v_proj):
    """utils/fused_utils.py:45-143 (same signature; `module` only names the device there)."""
    return fuse_linears([q_proj, k_proj, v_proj])


def _slices(sizes):
    out, lo = [], 0
    for s in sizes:
        out.append((lo, lo + s))
        lo += s
    return out


def _packed(*mods):
    return all(isinstance(m, WQLinear_GEMM) for m in mods) or all(isinstance(m, W8A8Linear) for m in mods)


def _same(mods):
    f = mods[0]
    if isinstance(f, W8A8Linear):
        return all(isinstance(m, W8A8Linear) and m.in_features == f.in_features and m.out_dtype == f.out_dtype and
                   ((m.smooth is None and f.smooth is None) or (m.smooth is not None and f.smooth is not None and torch.equal(m.smooth, f.smooth)))
                   for m in mods)
    return all(isinstance(m, WQLinear_GEMM) and (m.in_features, m.group_size, m.scales.dtype) == (f.in_features, f.group_size, f.scales.dtype)
               for m in mods)


def fuse_projections(denoiser):
    """Build the fused modules of a quantised skeleton (skeletons.UNet2DConditionSkeleton / MMDiTSkeleton).  They live in
    the owners' `__dict__` (not registered: the state dict and `get_named_linears` keep seeing the members only) and are
    picked up by the skeleton forwards.  Returns {kind: number of member Linears fused}.  Call again after the packed
    buffers were replaced (quantize(), load_state_dict())."""
    from . import skeletons as S
    unfuse_projections(denoiser)
    done = {"self_qkv": 0, "context_kv": 0, "time_emb_proj": 0, "adaln": 0, "geglu": 0}
    cross, resnets, adaln = [], [], []
    for m in denoiser.modules():
        if isinstance(m, S.Attention):
            if m.to_k.in_features == m.to_q.in_features and _packed(m.to_q, m.to_k, m.to_v) and _same([m.to_q, m.to_k, m.to_v]):
                m.__dict__["_qkv"] = fuse_qkv(m, m.to_q, m.to_k, m.to_v)
                done["self_qkv"] += 3
            elif m.to_k.in_features != m.to_q.in_features and _packed(m.to_k, m.to_v):
                cross.append(m)
        elif isinstance(m, S.JointAttention):
            if _packed(m.to_q, m.to_k, m.to_v) and _same([m.to_q, m.to_k, m.to_v]):
                m.__dict__["_qkv"] = fuse_qkv(m, m.to_q, m.to_k, m.to_v)
                done["self_qkv"] += 3
            if _packed(m.add_q_proj, m.add_k_proj, m.add_v_proj) and _same([m.add_q_proj, m.add_k_proj, m.add_v_proj]):
                m.__dict__["_add_qkv"] = fuse_qkv(m, m.add_q_proj, m.add_k_proj, m.add_v_proj)
                done["self_qkv"] += 3
        elif isinstance(m, S.GEGLU) and _packed(m.proj) and m.proj.out_features % 16 == 0:
            m.__dict__["_fused_act"] = True   # h * gelu(gate) by ops.geglu (one HBM pass)
            done["geglu"] += 1
        elif isinstance(m, S.ResnetBlock2D) and _packed(m.time_emb_proj):
            resnets.append(m)
        elif isinstance(m, (S.AdaLayerNormZero, S.AdaLayerNormContinuous)) and _packed(m.linear):
            adaln.append(m)
    if len(cross) > 0 and _same([l for a in cross for l in (a.to_k, a.to_v)]):
        mods = [l for a in cross for l in (a.to_k, a.to_v)]
        sl = _slices([l.out_features for l in mods])
        denoiser.__dict__["_ctx_kv"] = (fuse_linears(mods), [(a, sl[2 * i], sl[2 * i + 1]) for i, a in enumerate(cross)])
        done["context_kv"] = len(mods)
    if len(resnets) > 1 and _same([r.time_emb_proj for r in resnets]):
        mods = [r.time_emb_proj for r in resnets]
        denoiser.__dict__["_temb_all"] = (fuse_linears(mods), list(zip(resnets, _slices([l.out_features for l in mods]))))
        done["time_emb_proj"] = len(mods)
    if len(adaln) > 1 and _same([a.linear for a in adaln]):
        mods = [a.linear for a in adaln]
        denoiser.__dict__["_adaln_all"] = (fuse_linears(mods), list(zip(adaln, _slices([l.out_features for l in mods]))))
        done["adaln"] = len(mods)
    return done


def unfuse_projections(denoiser):
    for m in denoiser.modules():
        for key in ("_qkv", "_add_qkv", "_ctx_kv", "_temb_all", "_adaln_all", "_pre", "_fused_act"):
            m.__dict__.pop(key, None)


def grouped_prepass(denoiser, context=None, temb=None, emb=None):
    """Run the per-step grouped launches and hand every member its slice (a view of the one output):
    `_pre` on a cross-attention module = (k, v); on a ResNet block / AdaLN module = its projection of silu(temb / emb)."""
    ent = denoiser.__dict__.get("_ctx_kv")
    if ent is not None and context is not None:
        y = ent[0](context)
        for attn, (k0, k1), (v0, v1) in ent[1]:
            attn.__dict__["_pre"] = (y[..., k0:k1], y[..., v0:v1])
    ent = denoiser.__dict__.get("_temb_all")
    if ent is not None and temb is not None:
        y = ent[0](F.silu(temb))
        for res, (lo, hi) in ent[1]:
            res.__dict__["_pre"] = y[..., lo:hi]
    ent = denoiser.__dict__.get("_adaln_all")
    if ent is not None and emb is not None:
        y = ent[0](F.silu(emb))
        for mod, (lo, hi) in ent[1]:
            mod.__dict__["_pre"] = y[..., lo:hi]
