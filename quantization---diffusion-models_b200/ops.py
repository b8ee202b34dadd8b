"""Torch-tensor front end of the C ABI (include/qdm.h).

Every function takes CUDA tensors, allocates the outputs with torch (the library never
allocates), passes raw device pointers plus the current CUDA stream, and maps return codes
to the reference's exception types.  Nothing here computes anything in Python or torch:
a CPU tensor, a missing library or a non-B200 device is an error, not a fallback.
"""
import torch

from . import _lib
from ._lib import QDM_BF16, QDM_F16, QDM_F32, QDM_Q_NO_CLAMP, QDM_Q_ZERO_POINT, check

_DTYPES = {torch.float16: QDM_F16, torch.bfloat16: QDM_BF16, torch.float32: QDM_F32}
_workspaces = {}


def _dt(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise ValueError(f"unsupported dtype {t.dtype}; libqdm handles float16, bfloat16, float32")


def _cuda(t, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def _guard(device):
    """`with torch.cuda.device(d)` costs ~10 us per call; skip it when d is already current (the common case)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ws(device, nbytes):
    """Grow-only per-device scratch buffer handed to the library as its workspace."""
    key = (device.type, device.index)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def lib():
    return _lib.load()


def set_gemm_mode(ctas=0):
    """0 = heuristic, 1 = single-CTA tiles, 2 = CTA-pair (cta_group::2) tiles; bring-up / A-B timing only."""
    check(lib().qdm_set_gemm_mode(int(ctas)))


W4_NO_SMALLM, W4_NO_SKINNY, W4_NO_TMA, W4_NO_BSTAT, W4_NO_SK, W4_NO_RP = 1, 2, 4, 8, 16, 32


def set_w4_disable(mask=0):
    """OR-mask of W4_NO_* bits: takes kernel families out of the W4A16 dispatch for this process (tests / A-B timing);
    0 restores the heuristic.  The QDM_W4_NO_* environment switches are read once and cannot be flipped in-process."""
    check(lib().qdm_set_w4_disable(int(mask)))


def launch_count(reset=False):
    return int(lib().qdm_launch_count(1 if reset else 0))


# ------------------------------------------------------------------ (a) reductions
def _as_2d(x):
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    return x2


def colabsmax(x, out=None, running=False):
    """|x|.reshape(-1, C).amax(0) -> [C] in x.dtype (utils/calib_data.py:117-118,
    quantize/quantizer_SQ.py:417-418).  running=True folds into `out` (quantizer_SQ.py:1080-1084)."""
    _cuda(x, "x")
    x2 = _as_2d(x)
    rows, cols = x2.shape
    if out is None:
        if running:
            raise ValueError("running=True needs an existing `out`")
        out = torch.empty(cols, dtype=x.dtype, device=x.device)
    L = lib()
    ws = _ws(x.device, L.qdm_colreduce_workspace_bytes(rows, cols))
    with _guard(x.device):
        check(L.qdm_colabsmax(x2.data_ptr(), _dt(x2), rows, cols, x2.stride(0), out.data_ptr(),
                              1 if running else 0, ws.data_ptr(), ws.numel(), _stream(x)))
    return out


def colabssum(x):
    """sum over rows of |x| in fp32 with a fixed tree -> fp32 [C] (quantize/quantizer.py:652-657)."""
    _cuda(x, "x")
    x2 = _as_2d(x)
    rows, cols = x2.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    L = lib()
    ws = _ws(x.device, L.qdm_colreduce_workspace_bytes(rows, cols))
    with _guard(x.device):
        check(L.qdm_colabssum(x2.data_ptr(), _dt(x2), rows, cols, x2.stride(0), out.data_ptr(),
                              ws.data_ptr(), ws.numel(), _stream(x)))
    return out


def colstats(x, out_max=None, running=False, acc_maxsum=None, acc_abssum=None):
    """One read of x -> per-call column |x| max (into `out_max`, x.dtype; running=True folds max(out_max, .)),
    `acc_maxsum += colmax` and `acc_abssum += sum_rows |x|` (float64 [C], updated in place).  The fused form of
    the hook statistic (utils/calib_data.py:112-121 + StableDiffusion1_x.py:104-112) and of x_mean
    (quantize/quantizer.py:642-659).  Pass only the outputs you need."""
    _cuda(x, "x")
    x2 = _as_2d(x)
    rows, cols = x2.shape
    if running and out_max is None:
        raise ValueError("running=True needs an existing `out_max`")
    if out_max is None and acc_maxsum is None and acc_abssum is None:
        out_max = torch.empty(cols, dtype=x.dtype, device=x.device)
    for name, t, dt in (("out_max", out_max, x.dtype), ("acc_maxsum", acc_maxsum, torch.float64),
                        ("acc_abssum", acc_abssum, torch.float64)):
        if t is not None and (t.dtype != dt or t.numel() != cols or not t.is_contiguous() or t.device != x.device):
            raise ValueError(f"{name} must be a contiguous {dt} tensor of {cols} elements on {x.device}")
    L = lib()
    ws = _ws(x.device, L.qdm_colstats_workspace_bytes(rows, cols))
    with _guard(x.device):
        check(L.qdm_colstats(x2.data_ptr(), _dt(x2), rows, cols, x2.stride(0), _ptr(out_max), 1 if running else 0,
                             _ptr(acc_maxsum), _ptr(acc_abssum), ws.data_ptr(), ws.numel(), _stream(x)))
    return out_max


def rowabsmax(x):
    """x.abs().max(dim=-1) over contiguous rows -> [rows] (quantize/fake_quant.py:89,114)."""
    _cuda(x, "x")
    x2 = x.contiguous().reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    out = torch.empty(rows, dtype=x.dtype, device=x.device)
    with _guard(x.device):
        check(lib().qdm_rowabsmax(x2.data_ptr(), _dt(x2), rows, cols, out.data_ptr(), _stream(x)))
    return out.reshape(x.shape[:-1])


def absmax(x):
    """x.abs().max() -> 0-dim tensor (quantize/fake_quant.py:101,163)."""
    _cuda(x, "x")
    xc = x.contiguous()
    out = torch.empty((), dtype=x.dtype, device=x.device)
    L = lib()
    ws = _ws(x.device, L.qdm_absmax_workspace_bytes(xc.numel()))
    with _guard(x.device):
        check(L.qdm_absmax(xc.data_ptr(), _dt(xc), xc.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x)))
    return out


def awq_wsum(w, group):
    """sum_n |W|/(groupmax+1e-6) per input channel in fp32 (quantize/quantizer.py:627-637);
    the caller divides by N and casts.  W is [N, K] contiguous."""
    _cuda(w, "w")
    wc = w.contiguous()
    n, k = wc.shape
    out = torch.empty(k, dtype=torch.float32, device=w.device)
    L = lib()
    ws = _ws(w.device, L.qdm_colreduce_workspace_bytes(n, k))
    with _guard(w.device):
        check(L.qdm_awq_wsum(wc.data_ptr(), _dt(wc), n, k, int(group), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(w)))
    return out


def sqdiff_sum(a, b, out=None):
    """(a - b).float().pow(2).sum() as a float64 0-dim tensor (quantize/quantizer.py:777).  `out`: a one-element float64
    device tensor (e.g. a slot of the search's loss table) written in place."""
    _cuda(a, "a"), _cuda(b, "b")
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("sqdiff_sum: shape/dtype mismatch")
    ac, bc = a.contiguous(), b.contiguous()
    if out is None:
        out = torch.empty((), dtype=torch.float64, device=a.device)
    elif out.dtype != torch.float64 or out.numel() != 1 or out.device != a.device:
        raise ValueError("sqdiff_sum: out must be a one-element float64 tensor on the inputs' device")
    L = lib()
    ws = _ws(a.device, L.qdm_sqdiff_workspace_bytes(ac.numel()))
    with _guard(a.device):
        check(L.qdm_sqdiff_sum(ac.data_ptr(), bc.data_ptr(), _dt(ac), ac.numel(), out.data_ptr(),
                               ws.data_ptr(), ws.numel(), _stream(a)))
    return out


def awq_clip_search(w, x, group, n_bits=4, zero_point=True, n_grid=20, max_shrink=0.5):
    """AwqQuantizer._compute_best_clip (quantize/quantizer.py:805-863) for one Linear: w [co, ci], x [n_tok, ci] (the
    already subsampled tokens; a strided view is fine) -> best_max [co, ci / group, 1] in w.dtype.  One Gram pass over x
    + one search kernel (include/qdm.h: qdm_awq_clip_search)."""
    _cuda(w, "w"), _cuda(x, "x")
    if w.dim() != 2 or x.dim() != 2 or x.shape[1] != w.shape[1] or x.dtype != w.dtype or x.stride(1) != 1:
        raise ValueError(f"awq_clip_search: w{tuple(w.shape)} / x{tuple(x.shape)} must be 2-D with the same K and dtype")
    wc = w.contiguous()
    co, ci = wc.shape
    out = torch.empty((co, ci // group), dtype=w.dtype, device=w.device)
    L = lib()
    ws = _ws(w.device, L.qdm_awq_clip_workspace_bytes(ci, int(group)))
    with _guard(w.device):
        check(L.qdm_awq_clip_search(wc.data_ptr(), _dt(wc), co, ci, int(group), int(n_bits), QDM_Q_ZERO_POINT if zero_point else 0,
                                    x.data_ptr(), x.shape[0], x.stride(0), int(n_grid), float(max_shrink), out.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream(w)))
    return out.unsqueeze(-1)


# ------------------------------------------------------------------ (b) quantise / pack
def _flags(zero_point, no_clamp):
    if zero_point and no_clamp:
        raise ValueError("zero_point and no_clamp are mutually exclusive")
    return (QDM_Q_ZERO_POINT if zero_point else 0) | (QDM_Q_NO_CLAMP if no_clamp else 0)


def quant_group(w, group, n_bits=4, zero_point=True, no_clamp=False, pre_mul=None, clip_max=None,
                post_div=None, want_dq=True, want_codes=False, want_scales=True, out=None):
    """Per-group RTN over the last dim (see qdm_quant_group in include/qdm.h).
    Returns (dq, codes, scales, zeros); entries not requested are None.  `out` may alias `w`."""
    _cuda(w, "w")
    wc = w.contiguous()
    k = wc.shape[-1]
    n = wc.numel() // k
    if group <= 0:
        group = k
    if k % group:
        raise ValueError(f"group {group} must divide the last dim {k}")
    G = k // group
    dev, dt = w.device, w.dtype
    for name, v in (("pre_mul", pre_mul), ("post_div", post_div)):
        if v is not None and (v.numel() != k or v.dtype != dt):
            raise ValueError(f"{name} must be a [{k}] tensor of dtype {dt}")
    if clip_max is not None and (clip_max.numel() != n * G or clip_max.dtype != dt):
        raise ValueError(f"clip_max must hold {n * G} values of dtype {dt}")
    dq = (out if out is not None else torch.empty_like(wc)) if want_dq else None
    codes = torch.empty(wc.shape, dtype=torch.uint8 if zero_point else torch.int8, device=dev) if want_codes else None
    scales = torch.empty((n, G), dtype=dt, device=dev) if want_scales else None
    zeros = torch.empty((n, G), dtype=dt, device=dev) if (want_scales and zero_point) else None
    pm = pre_mul.contiguous() if pre_mul is not None else None
    pd = post_div.contiguous() if post_div is not None else None
    cm = clip_max.contiguous() if clip_max is not None else None
    with _guard(dev):
        check(lib().qdm_quant_group(wc.data_ptr(), _dt(wc), n, k, int(group), int(n_bits), _flags(zero_point, no_clamp),
                                    _ptr(pm), _ptr(cm), _ptr(pd), _ptr(dq), _ptr(codes), _ptr(scales), _ptr(zeros),
                                    _stream(w)))
    return dq, codes, scales, zeros


def quant_rowwise(x, n_bits=8, zero_point=False, no_clamp=True, want_dq=True, want_codes=False, want_scales=False):
    """One scale per row of the last dim (fake_quant.py:86-93,109-118)."""
    _cuda(x, "x")
    xc = x.contiguous()
    cols = xc.shape[-1]
    rows = xc.numel() // cols
    dq = torch.empty_like(xc) if want_dq else None
    codes = torch.empty(xc.shape, dtype=torch.uint8 if zero_point else torch.int8, device=x.device) if want_codes else None
    scales = torch.empty(rows, dtype=x.dtype, device=x.device) if want_scales else None
    zeros = torch.empty(rows, dtype=x.dtype, device=x.device) if (want_scales and zero_point) else None
    with _guard(x.device):
        check(lib().qdm_quant_rowwise(xc.data_ptr(), _dt(xc), rows, cols, int(n_bits), _flags(zero_point, no_clamp),
                                      _ptr(dq), _ptr(codes), _ptr(scales), _ptr(zeros), _stream(x)))
    return dq, codes, scales, zeros


def quant_tensor(x, n_bits=8, want_dq=True, want_codes=False):
    """Whole-tensor absmax RTN (fake_quant.py:97-105,158-167). Returns (dq, codes, scale)."""
    _cuda(x, "x")
    xc = x.contiguous()
    dq = torch.empty_like(xc) if want_dq else None
    codes = torch.empty(xc.shape, dtype=torch.int8, device=x.device) if want_codes else None
    scale = torch.empty((), dtype=x.dtype, device=x.device)
    L = lib()
    ws = _ws(x.device, L.qdm_quant_tensor_workspace_bytes(xc.numel()))
    with _guard(x.device):
        check(L.qdm_quant_tensor(xc.data_ptr(), _dt(xc), xc.numel(), int(n_bits), _ptr(dq), _ptr(codes),
                                 scale.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x)))
    return dq, codes, scale


def actquant_token_i8(x, smooth=None):
    """Per-token int8 codes + fp32 scales of x[..., K] -> (xq int8 [M, K], sx fp32 [M])."""
    _cuda(x, "x")
    x2 = x.contiguous().reshape(-1, x.shape[-1])
    rows, cols = x2.shape
    if smooth is not None and (smooth.numel() != cols or smooth.dtype != x.dtype):
        raise ValueError(f"smooth must be a [{cols}] tensor of dtype {x.dtype}")
    xq = torch.empty((rows, cols), dtype=torch.int8, device=x.device)
    sx = torch.empty(rows, dtype=torch.float32, device=x.device)
    sm = smooth.contiguous() if smooth is not None else None
    with _guard(x.device):
        check(lib().qdm_actquant_token_i8(x2.data_ptr(), _dt(x2), rows, cols, _ptr(sm), xq.data_ptr(), sx.data_ptr(), _stream(x)))
    return xq, sx


def pack_awq(codes_nk):
    """int codes [N, K] (nn.Linear orientation, low 4 bits) -> qweight int32 [K, N/8] in AWQ order."""
    _cuda(codes_nk, "codes")
    c = codes_nk.contiguous()
    if c.dtype not in (torch.int8, torch.uint8):
        raise ValueError("codes must be int8/uint8")
    n, k = c.shape
    if n % 8:
        raise ValueError(f"N={n} must be a multiple of 8")
    q = torch.empty((k, n // 8), dtype=torch.int32, device=c.device)
    with _guard(c.device):
        check(lib().qdm_pack_awq(c.data_ptr(), n, k, q.data_ptr(), _stream(c)))
    return q


def unpack_awq(qweight):
    """qweight int32 [K, N/8] -> codes int8 [K, N] in natural column order."""
    _cuda(qweight, "qweight")
    q = qweight.contiguous()
    k, nw = q.shape
    out = torch.empty((k, nw * 8), dtype=torch.int8, device=q.device)
    with _guard(q.device):
        check(lib().qdm_unpack_awq(q.data_ptr(), k, nw * 8, out.data_ptr(), _stream(q)))
    return out


def quant_pack_awq(w, group, want_dq=False):
    """W [N, K] -> (qweight [K, N/8] int32, qzeros [K/g, N/8] int32, scales [K/g, N] dtype, dq|None)."""
    _cuda(w, "w")
    wc = w.contiguous()
    n, k = wc.shape
    if group <= 0 or k % group:
        raise ValueError(f"group {group} must divide K={k}")
    if n % 8:
        raise ValueError(f"N={n} must be a multiple of 8")
    if group not in (32, 64, 128, 256) or (w.dtype == torch.float32 and group > 128):
        # shapes the fused kernel does not tile: same result from the unfused kernels
        dq, codes, s, z = quant_group(wc, group, 4, zero_point=True, want_dq=want_dq, want_codes=True)
        return pack_awq(codes), pack_awq(z.to(torch.int8)), s.t().contiguous(), dq
    qweight = torch.empty((k, n // 8), dtype=torch.int32, device=w.device)
    qzeros = torch.empty((k // group, n // 8), dtype=torch.int32, device=w.device)
    scales = torch.empty((k // group, n), dtype=w.dtype, device=w.device)
    dq = torch.empty_like(wc) if want_dq else None
    with _guard(w.device):
        check(lib().qdm_quant_pack_awq(wc.data_ptr(), _dt(wc), n, k, int(group), qweight.data_ptr(), qzeros.data_ptr(),
                                       scales.data_ptr(), _ptr(dq), _stream(w)))
    return qweight, qzeros, scales, dq


def dequant_awq(qweight, qzeros, scales, group):
    """(q - z) * s -> W_kn [K, N] in scales.dtype (utils/packing_utils.py:87-102)."""
    _cuda(qweight, "qweight")
    k, nw = qweight.shape
    out = torch.empty((k, nw * 8), dtype=scales.dtype, device=qweight.device)
    with _guard(qweight.device):
        check(lib().qdm_dequant_awq(qweight.contiguous().data_ptr(), qzeros.contiguous().data_ptr(),
                                    scales.contiguous().data_ptr(), _dt(scales), k, nw * 8, int(group),
                                    out.data_ptr(), _stream(qweight)))
    return out


def geglu(x):
    """h * F.gelu(gate) with (h, gate) = x.chunk(2, -1): the activation between ff.net.0.proj and ff.net.2 of a diffusers
    FeedForward (the reference quantizes both Linears, models/StableDiffusion1_x.py:121-137), one HBM pass."""
    _cuda(x, "x")
    f2 = x.shape[-1]
    if f2 % 16:
        raise ValueError(f"geglu: last dimension {f2} must be a multiple of 16")
    x2 = x.reshape(-1, f2)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    y = torch.empty((*x.shape[:-1], f2 // 2), dtype=x.dtype, device=x.device)
    with _guard(x.device):
        check(lib().qdm_geglu(x2.data_ptr(), _dt(x), x2.shape[0], f2 // 2, y.data_ptr(), _stream(x)))
    return y


# ------------------------------------------------------------------ (c)(d) GEMMs
def _gemm_io(x, n_out, out_dtype=None):
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    y = torch.empty((x2.shape[0], n_out), dtype=out_dtype or x.dtype, device=x.device)
    return x2, y


def gemm_f16(x, w, bias=None):
    """F.linear(x, w, bias) with w [N, K] (fake-quant weights, quantize/fake_quant.py:223)."""
    _cuda(x, "x"), _cuda(w, "w")
    if w.dtype != x.dtype:
        raise ValueError("x and w must share a dtype")
    if w.dim() != 2 or x.shape[-1] != w.shape[1]:
        raise ValueError(f"shape mismatch: x[..., {x.shape[-1]}] @ w{tuple(w.shape)}^T")
    x2, y = _gemm_io(x, w.shape[0])
    wc = w.contiguous()
    b = bias.to(x.dtype).contiguous() if bias is not None else None
    with _guard(x.device):
        check(lib().qdm_gemm_f16(x2.data_ptr(), wc.data_ptr(), _ptr(b), y.data_ptr(), _dt(x2),
                                 x2.shape[0], wc.shape[0], wc.shape[1], _stream(x)))
    return y.reshape(*x.shape[:-1], w.shape[0])


def gemm_f16_kn(x, w_kn, bias=None):
    """x @ w_kn + bias with w_kn [K, N]."""
    _cuda(x, "x"), _cuda(w_kn, "w_kn")
    if w_kn.dtype != x.dtype:
        raise ValueError("x and w must share a dtype")
    if w_kn.dim() != 2 or x.shape[-1] != w_kn.shape[0]:
        raise ValueError(f"shape mismatch: x[..., {x.shape[-1]}] @ w_kn{tuple(w_kn.shape)}")
    x2, y = _gemm_io(x, w_kn.shape[1])
    wc = w_kn.contiguous()
    b = bias.to(x.dtype).contiguous() if bias is not None else None
    with _guard(x.device):
        check(lib().qdm_gemm_f16_kn(x2.data_ptr(), wc.data_ptr(), _ptr(b), y.data_ptr(), _dt(x2),
                                    x2.shape[0], wc.shape[1], wc.shape[0], _stream(x)))
    return y.reshape(*x.shape[:-1], w_kn.shape[1])


_gemm_ws = {}   # device index -> stream-K workspace tensor (kept alive for the life of the process)


def _set_gemm_workspace(L, device):
    """One stream-K workspace per device (include/qdm.h: qdm_gemm_set_workspace); GEMM calls of a device are expected on
    one stream at a time, as in the reference's single-stream execution."""
    buf = torch.empty(int(L.qdm_gemm_workspace_bytes()), dtype=torch.uint8, device=device)
    with _guard(device):
        check(L.qdm_gemm_set_workspace(buf.data_ptr(), buf.numel(), torch.cuda.current_stream(device).cuda_stream))
    _gemm_ws[device.index] = buf


GEMM_VARIANTS = {1: "single", 2: "pair", 3: "bstat", 4: "streamk", 5: "skinny", 6: "smallm", 7: "rp1", 8: "rp2", 9: "quad", 10: "ts"}


def gemm_last_variant():
    """(name, tile width) of the kernel the last W4A16 GEMM call of this thread launched (tests assert the dispatch)."""
    import ctypes
    tile = ctypes.c_int(0)
    v = lib().qdm_gemm_last_variant(ctypes.byref(tile))
    return GEMM_VARIANTS.get(int(v), str(v)), int(tile.value)


def w4a16_repack(qweight, qzeros, scales, group):
    """Kernel-native copy of an AWQ weight (include/qdm.h: qdm_w4a16_repack) -> uint8 blob for gemm_w4a16(..., blob=)."""
    _cuda(qweight, "qweight")
    k, nw = qweight.shape
    n = nw * 8
    if tuple(qzeros.shape) != (k // group, nw) or tuple(scales.shape) != (k // group, n):
        raise ValueError(f"qzeros{tuple(qzeros.shape)} / scales{tuple(scales.shape)} do not match qweight{tuple(qweight.shape)}, group {group}")
    L = lib()
    nbytes = int(L.qdm_w4a16_repack_bytes(n, k))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=qweight.device)
    with _guard(qweight.device):
        check(L.qdm_w4a16_repack(qweight.contiguous().data_ptr(), qzeros.contiguous().data_ptr(), scales.contiguous().data_ptr(),
                                 n, k, int(group), blob.data_ptr(), nbytes, _stream(qweight)))
    return blob


def w4a16_repack_ts(qweight, qzeros, scales, group):
    """Kernel-native copy of an AWQ weight for the TMEM-A kernel (include/qdm.h: qdm_w4a16_repack_ts) -> uint8 blob."""
    _cuda(qweight, "qweight")
    k, nw = qweight.shape
    n = nw * 8
    if tuple(qzeros.shape) != (k // group, nw) or tuple(scales.shape) != (k // group, n):
        raise ValueError(f"qzeros{tuple(qzeros.shape)} / scales{tuple(scales.shape)} do not match qweight{tuple(qweight.shape)}, group {group}")
    L = lib()
    nbytes = int(L.qdm_w4a16_repack_ts_bytes(n, k))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=qweight.device)
    with _guard(qweight.device):
        check(L.qdm_w4a16_repack_ts(qweight.contiguous().data_ptr(), qzeros.contiguous().data_ptr(), scales.contiguous().data_ptr(),
                                    _dt(scales), n, k, int(group), blob.data_ptr(), nbytes, _stream(qweight)))
    return blob


def gemm_w4a16(x, qweight, qzeros, scales, group, bias=None, blob=None, blob_ts=None, out=None):
    """x @ dequant(qweight, qzeros, scales) + bias, AWQ GEMM layout (quantize/quantizer.py:544-569).
    `blob` = w4a16_repack(...) of the same weight routes M > 128 problems to the repacked-weight kernel.
    This is the per-Linear hot call of a denoise step: the Python side is kept to the bare minimum."""
    if not x.is_cuda or not qweight.is_cuda:
        raise RuntimeError("gemm_w4a16 needs CUDA tensors: the B200 path has no CPU fallback")
    if scales.dtype != x.dtype:
        raise ValueError("scales must have the activation dtype")
    k, nw = qweight.shape
    n = nw * 8
    if x.shape[-1] != k:
        raise ValueError(f"x has K={x.shape[-1]}, qweight has K={k}")
    x2 = x.reshape(-1, k)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    m = x2.shape[0]
    if out is None:
        y = torch.empty((m, n), dtype=x.dtype, device=x.device)
    else:   # caller-owned output (a [M, N] contiguous view): no allocation in the call
        if out.shape != (m, n) or out.dtype != x.dtype or not out.is_contiguous() or out.device != x.device:
            raise ValueError(f"out must be a contiguous [{m}, {n}] {x.dtype} tensor on {x.device}")
        y = out
    if bias is not None and (bias.dtype != x.dtype or not bias.is_contiguous()):
        bias = bias.to(x.dtype).contiguous()
    L = _lib._lib or lib()
    if x.device.index not in _gemm_ws:
        _set_gemm_workspace(L, x.device)
    with _guard(x.device):
        rc = L.qdm_gemm_w4a16_plan(x2.data_ptr(), qweight.data_ptr(), qzeros.data_ptr(), scales.data_ptr(),
                                   0 if blob is None else blob.data_ptr(), 0 if blob_ts is None else blob_ts.data_ptr(),
                                   0 if bias is None else bias.data_ptr(), y.data_ptr(),
                                   _DTYPES[x.dtype], m, n, k, group, torch.cuda.current_stream(x.device).cuda_stream)
    if rc:
        check(rc)
    return y.reshape(*x.shape[:-1], n) if x.dim() != 2 else y


def gemm_w8a8(xq, sx, wq, sw, bias=None, out_dtype=torch.float16):
    """(xq @ wq.T) * sx[:, None] * sw[None, :] + bias with int8 codes and fp32 scales."""
    _cuda(xq, "xq"), _cuda(wq, "wq")
    m, k = xq.shape
    n = wq.shape[0]
    if wq.shape[1] != k or sx.numel() != m or sw.numel() != n or sx.dtype != torch.float32 or sw.dtype != torch.float32:
        raise ValueError(f"shape/dtype mismatch: xq{tuple(xq.shape)} wq{tuple(wq.shape)} sx{tuple(sx.shape)} sw{tuple(sw.shape)}")
    y = torch.empty((m, n), dtype=out_dtype, device=xq.device)
    b = bias.to(out_dtype).contiguous() if bias is not None else None
    with _guard(xq.device):
        check(lib().qdm_gemm_w8a8(xq.data_ptr(), sx.data_ptr(), wq.data_ptr(), sw.data_ptr(), _ptr(b), y.data_ptr(),
                                  _DTYPES[out_dtype], m, n, k, _stream(xq)))
    return y


# ------------------------------------------------------------------ 3x3 convolution as an implicit GEMM
def conv3x3_weight_taps(w):
    """[N, C, 3, 3] -> [N, 9*C] in (n, dy, dx, c) order: the K axis of the implicit GEMM (tap-major, channels inside)."""
    n, c, kh, kw = w.shape
    if (kh, kw) != (3, 3):
        raise ValueError(f"expected a 3x3 kernel, got {kh}x{kw}")
    return w.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous()


def conv3x3_stride2_ok(h, w):
    """True when the 3x3 / stride 2 / padding 1 convolution of an h x w grid runs on `qdm_conv3x3s2_*` (even sizes whose
    output grid tiles by whole image rows, see include/qdm.h)."""
    return h % 2 == 0 and w % 2 == 0 and bool(lib().qdm_conv3x3_direct_ok(h // 2, w // 2))


def _conv3x3_io(x, n_out, padded=None, stride=1):
    """Returns (x_in, y_out, (b, h, w, c), direct).  direct (W divides 128, see qdm_conv3x3_direct_ok): x as unpadded
    NHWC -- a free view when x is channels-last in memory, one transpose copy otherwise -- and a dense NHWC output.
    Otherwise the padded-grid form: NHWC with a one-pixel zero border and an output on the same grid.  stride 2 exists
    in the direct form only (output [B, H/2, W/2, N])."""
    _cuda(x, "x")
    if x.dim() != 4:
        raise ValueError(f"expected [B, C, H, W], got {tuple(x.shape)}")
    b, c, h, w = x.shape
    if stride == 2:
        if padded:
            raise ValueError("the stride-2 convolution has no padded-grid form")
        if not conv3x3_stride2_ok(h, w):
            raise ValueError(f"stride-2 convolution of a {h} x {w} grid is not supported (even sizes, output width dividing 128)")
        return x.permute(0, 2, 3, 1).contiguous(), torch.empty((b, h // 2, w // 2, n_out), dtype=x.dtype, device=x.device), (b, h, w, c), True
    if stride != 1:
        raise ValueError(f"stride must be 1 or 2, got {stride}")
    # measured (profiles/conv3x3_r01.json): the direct form wins for rows of >= 32 pixels (no padded copy, no border
    # rows); for 16- and 8-pixel rows the 4-D boxes are fetched more slowly than they save and the padded grid wins
    direct = (w >= 32 and bool(lib().qdm_conv3x3_direct_ok(h, w))) if padded is None else not padded
    if direct:
        x_in = x.permute(0, 2, 3, 1).contiguous()
        y = torch.empty((b, h, w, n_out), dtype=x.dtype, device=x.device)
    else:
        x_in = torch.nn.functional.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()
        y = torch.empty((b, h + 2, w + 2, n_out), dtype=x.dtype, device=x.device)
    return x_in, y, (b, h, w, c), direct


def _conv3x3_out(y, h, w, direct):
    """logical [B, N, H, W] view of the NHWC result (channels-last memory format; dense in the direct form, the
    interior of the padded grid otherwise)"""
    return (y if direct else y[:, 1:h + 1, 1:w + 1, :]).permute(0, 3, 1, 2)


def conv3x3_f16(x, w_tap, bias=None, padded=None, stride=1):
    """F.conv2d(x, w, bias, stride=stride, padding=1) for a 3x3 kernel, w_tap = conv3x3_weight_taps(w) (fake-quant weights,
    quantize/fake_quant.py:337-341), as one tcgen05 GEMM whose A rows are fetched per tap by TMA; C % 64 == 0,
    N % 8 == 0.  `padded` forces the padded-grid (True) or direct (False) form; default: direct when the geometry allows.
    stride 2 (the down-samplers): direct form only, see conv3x3_stride2_ok."""
    _cuda(w_tap, "w_tap")
    if w_tap.dtype != x.dtype or w_tap.dim() != 2 or w_tap.shape[1] != 9 * x.shape[1]:
        raise ValueError(f"w_tap must be [N, {9 * x.shape[1]}] of dtype {x.dtype}")
    x_in, y, (b, h, w, c), direct = _conv3x3_io(x, w_tap.shape[0], padded, stride)
    wt = w_tap.contiguous()
    bs = bias.to(x.dtype).contiguous() if bias is not None else None
    fn = lib().qdm_conv3x3s2_nhwc_f16 if stride == 2 else lib().qdm_conv3x3_nhwc_f16 if direct else lib().qdm_conv3x3_f16
    with _guard(x.device):
        check(fn(x_in.data_ptr(), wt.data_ptr(), _ptr(bs), y.data_ptr(), _dt(x_in), b, h, w, c, wt.shape[0], _stream(x)))
    return _conv3x3_out(y, h, w, direct)


def conv3x3_w4a16(x, qweight, qzeros, scales, group, bias=None, padded=None, stride=1):
    """The same convolution from AWQ-packed int4 weights of w_tap (qweight [9C, N/8], qzeros / scales per group)."""
    _cuda(qweight, "qweight")
    n = scales.shape[1]
    if scales.dtype != x.dtype:
        raise ValueError("x and scales must share a dtype")
    if qweight.shape[0] != 9 * x.shape[1] or qweight.shape[1] * 8 != n:
        raise ValueError(f"qweight must be [{9 * x.shape[1]}, N/8]")
    x_in, y, (b, h, w, c), direct = _conv3x3_io(x, n, padded, stride)
    bs = bias.to(x.dtype).contiguous() if bias is not None else None
    fn = lib().qdm_conv3x3s2_nhwc_w4a16 if stride == 2 else lib().qdm_conv3x3_nhwc_w4a16 if direct else lib().qdm_conv3x3_w4a16
    with _guard(x.device):
        check(fn(x_in.data_ptr(), qweight.data_ptr(), qzeros.data_ptr(), scales.data_ptr(), _ptr(bs), y.data_ptr(),
                 _dt(x_in), b, h, w, c, n, int(group), _stream(x)))
    return _conv3x3_out(y, h, w, direct)
