"""GPU parity for kernels (c) W4A16 and (d) W8A8 (tcgen05/TMEM GEMMs) through the C ABI.

Tolerance (north star): max |y - ref| / max |ref| <= 1e-2 with fp32 accumulation.  The reference
output is the reference's own formulation on the same inputs: F.linear(x, W_fakequant, bias)
(quantize/fake_quant.py:223) evaluated in fp32 on the dequantised weights."""
import pytest
import torch

from _util import DT, max_rel_err

import oracle.qdm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2


def ref_linear(x, w_nk, bias=None):
    y = x.float().cuda() @ w_nk.float().cuda().t()
    if bias is not None:
        y = y + bias.float().cuda()
    return y.cpu()


SHAPES = [(128, 128, 64), (256, 256, 128), (300, 320, 320), (1024, 2432, 2432), (77, 1280, 768), (4096, 640, 2560),
          (1, 512, 256), (513, 72, 192)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_f16(qdm, dt, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DT[dt])
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    y = qdm.ops.gemm_f16(x.to(DEV), w.to(DEV), b.to(DEV))
    assert max_rel_err(y, ref_linear(x, w, b)) <= TOL
    y2 = qdm.ops.gemm_f16_kn(x.to(DEV), w.t().contiguous().to(DEV), None)
    assert max_rel_err(y2, ref_linear(x, w)) <= TOL


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K,group", [(128, 128, 128, 128), (256, 256, 256, 64), (300, 320, 320, 64),
                                          (1024, 2432, 2432, 128), (77, 1280, 768, 128), (4096, 2560, 320, 64),
                                          (1, 1024, 256, 128), (513, 72, 192, 64), (2048, 64, 2432, 128),
                                          (16, 1280, 1280, 128), (64, 320, 1280, 128), (33, 72, 192, 64), (2, 14592, 2432, 128), (48, 640, 320, 64),
                                          (60000, 320, 320, 64), (30001, 2560, 320, 64), (40960, 416, 384, 128), (57344, 160, 64, 64),
                                          # M <= 32 with more than 2 M weights: the sector-wide cluster-split-K kernel
                                          (8, 1280, 2816, 128), (1, 14592, 2432, 128), (3, 2432, 9728, 128)])
def test_gemm_w4a16(qdm, dt, M, N, K, group):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DT[dt])
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    if N % 64 == 0:
        qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w.to(DEV), group, want_dq=True)
        dq = dq.cpu()
    else:
        oq, oz, os_, dq = O.awq_from_linear(w, group, 4)
        qweight, qzeros, scales = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV)
    y = qdm.ops.gemm_w4a16(x.to(DEV), qweight, qzeros, scales, group, b.to(DEV))
    assert y.shape == (M, N) and y.dtype == DT[dt]
    assert max_rel_err(y, ref_linear(x, dq, b)) <= TOL
    # same product through the decoded weight: the in-mainloop dequant is bit-identical to
    # dequantize_gemm, so the two kernels may differ only by accumulation order
    y_kn = qdm.ops.gemm_f16_kn(x.to(DEV), qdm.ops.dequant_awq(qweight, qzeros, scales, group), b.to(DEV))
    # (one ulp of the output dtype at the largest magnitude: fp16 2^-11, bf16 2^-8)
    assert max_rel_err(y, y_kn.cpu()) <= (2e-3 if dt == "f16" else 8e-3)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 320, 320), (1024, 1280, 1280), (4096, 5120, 640), (77, 1280, 2048),
                                   (1, 256, 512), (200, 72, 144)])
def test_gemm_w8a8(qdm, dt, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    x[:, 3] *= 20
    x = x.to(DT[dt])
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    xq, sx = qdm.ops.actquant_token_i8(x.to(DEV))
    _, wq, sw, _ = qdm.ops.quant_rowwise(w.to(DEV), 8, want_dq=False, want_codes=True, want_scales=True)
    y = qdm.ops.gemm_w8a8(xq, sx, wq, sw.float(), b.to(DEV), out_dtype=DT[dt])
    # exact integer reference of the same codes
    yi = (xq.cpu().double() @ wq.cpu().double().t()) * sx.cpu().double()[:, None] * sw.cpu().double()[None, :] + b.double()
    assert max_rel_err(y, yi) <= 4e-3          # one rounding to the output dtype (bf16: 2^-8)
    # the reference's fake-quant formulation (fake_quant.py:86-93,109-118,223)
    assert max_rel_err(y, O.linear_w8a8_fake(x, w, b)) <= TOL


def test_gemm_large_m_round_trip(qdm):
    """BASELINE-size shape (SD1.5: M=65536, K=320) through a size-independent property: linearity.
    y(x1 + x2) == y(x1) + y(x2) up to output rounding, and row permutation equivariance."""
    M, N, K, group = 65536, 320, 320, 64
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
    qweight, qzeros, scales, _ = qdm.ops.quant_pack_awq(w, group)
    x1 = torch.randn(M, K, generator=g).half().to(DEV)
    perm = torch.randperm(M, generator=g).to(DEV)
    y1 = qdm.ops.gemm_w4a16(x1, qweight, qzeros, scales, group)
    yp = qdm.ops.gemm_w4a16(x1[perm].contiguous(), qweight, qzeros, scales, group)
    assert torch.equal(yp, y1[perm])
    y2 = qdm.ops.gemm_w4a16((x1 * 2).contiguous(), qweight, qzeros, scales, group)
    normal = y1.float().abs() >= 2.0 ** -13          # scaling by 2 is exact unless the fp16 output is subnormal
    assert torch.equal(y2.float()[normal], (y1.float() * 2)[normal])


def test_gemm_w4a16_stream_k(qdm):
    """Stream-K path (partial accumulators parked in the workspace, added back in slot order): forced on shapes whose
    tiles are split between CTA pairs in every way (two-way and many-way splits, odd K tail, ragged M / N), called twice
    so that the reader-reset flags are exercised; the result must equal the whole-tile kernel's within one output ulp and
    be identical from call to call (deterministic reduction order)."""
    g = torch.Generator().manual_seed(3)
    try:
        for M, N, K, group in [(4096, 1280, 1280, 128), (1232, 1280, 768, 128), (1024, 1280, 5120, 128), (513, 2560, 320, 64),
                               (777, 96, 192, 64), (4096, 1280, 5120, 128)]:
            x = torch.randn(M, K, generator=g).half().to(DEV)
            w = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
            b = torch.randn(N, generator=g).half().to(DEV)
            qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w, group, want_dq=True)
            qdm.ops.set_gemm_mode(2)
            y_tiles = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)
            qdm.ops.set_gemm_mode(8)
            y1 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)
            y2 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)
            assert torch.equal(y1, y2), (M, N, K)
            assert max_rel_err(y1, ref_linear(x.cpu(), dq.cpu(), b.cpu())) <= TOL
            assert max_rel_err(y1, y_tiles.cpu()) <= 2e-3
    finally:
        qdm.ops.set_gemm_mode(0)


# ------------------------------------------------------------------ 3x3 convolution as an implicit GEMM
def ref_conv3x3(x, w, bias, stride=1):
    """F.conv2d in fp32 without TF32, on the GPU (the reference's own op, fake_quant.py:339, at full precision)."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = torch.nn.functional.conv2d(x.float().cuda(), w.float().cuda(), None if bias is None else bias.float().cuda(), stride, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    return y.cpu()


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("B,C,N,H,W", [(2, 64, 64, 8, 8), (2, 320, 320, 16, 16), (1, 128, 320, 5, 7), (3, 640, 1280, 8, 8),
                                       (1, 64, 72, 3, 3), (4, 320, 640, 32, 32), (16, 320, 320, 64, 64), (1, 64, 64, 128, 128),
                                       (5, 128, 64, 4, 4), (2, 64, 64, 16, 8), (1, 64, 128, 12, 16)])
def test_conv3x3_implicit_gemm(qdm, dt, B, C, N, H, W):
    """qdm_conv3x3_f16 / qdm_conv3x3_w4a16 against F.conv2d(stride 1, padding 1) on the same (fake-quant) weights;
    the last case is the full-size SD1.5 resnet convolution (batch 8 + CFG, 64 x 64 latents)."""
    g = torch.Generator().manual_seed(B + C + N + H + W)
    x = torch.randn(B, C, H, W, generator=g).to(DT[dt])
    w = (torch.randn(N, C, 3, 3, generator=g) * 0.03).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    taps = qdm.ops.conv3x3_weight_taps(w.to(DEV))
    assert taps.shape == (N, 9 * C) and torch.equal(taps.cpu().reshape(N, 3, 3, C).permute(0, 3, 1, 2), w)
    ref = ref_conv3x3(x, w, b)
    direct_ok = bool(qdm.ops.lib().qdm_conv3x3_direct_ok(H, W))
    assert direct_ok == (128 % W == 0 and ((128 // W) % H == 0 or H % (128 // W) == 0))
    for xin in (x.to(DEV), x.to(DEV).contiguous(memory_format=torch.channels_last)):
        for padded in ([True, False] if direct_ok else [True]):      # padded-grid form and direct 4-D TMA form
            y = qdm.ops.conv3x3_f16(xin, taps, b.to(DEV), padded=padded)
            assert y.shape == (B, N, H, W) and y.dtype == DT[dt]
            assert max_rel_err(y, ref) <= TOL, (padded,)
            if not padded:
                assert y.is_contiguous(memory_format=torch.channels_last)
    if not direct_ok:
        with pytest.raises(RuntimeError, match="whole image rows"):
            qdm.ops.conv3x3_f16(x.to(DEV), taps, None, padded=False)
    assert max_rel_err(qdm.ops.conv3x3_f16(x.to(DEV), taps, None), ref_conv3x3(x, w, None)) <= TOL
    # packed int4 weights of the tap-major matrix: codes bit-exact vs the oracle's RTN, conv within tolerance
    group = 64
    qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(taps, group, want_dq=True) if N % 64 == 0 else (None,) * 4
    if qweight is None:
        oq, oz, os_, dq = O.awq_from_linear(taps.cpu(), group, 4)
        qweight, qzeros, scales = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV)
    else:
        assert torch.equal(dq.cpu(), O.rtn_group(taps.cpu(), group, True, 4)[0])
        dq = dq.cpu()
    w_dq = dq.reshape(N, 3, 3, C).permute(0, 3, 1, 2)
    ref4 = ref_conv3x3(x, w_dq, b)
    for padded in ([True, False] if direct_ok else [True]):
        y4 = qdm.ops.conv3x3_w4a16(x.to(DEV), qweight, qzeros, scales, group, b.to(DEV), padded=padded)
        assert y4.shape == (B, N, H, W) and y4.dtype == DT[dt]
        assert max_rel_err(y4, ref4) <= TOL, (padded,)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("B,C,N,H,W", [(16, 320, 320, 64, 64), (16, 640, 640, 32, 32), (16, 1280, 1280, 16, 16), (1, 64, 64, 16, 16),
                                       (3, 128, 72, 8, 8), (2, 64, 128, 32, 16), (1, 64, 64, 256, 256), (5, 64, 64, 4, 4),
                                       (2, 64, 64, 2, 2)])
def test_conv3x3_stride2_implicit_gemm(qdm, dt, B, C, N, H, W):
    """qdm_conv3x3s2_nhwc_f16 / _w4a16 (the UNet down-samplers: 3x3, stride 2, padding 1) against F.conv2d on the same
    (fake-quant) weights; the first three cases are the SD1.5 down-samplers at batch 8 + CFG."""
    g = torch.Generator().manual_seed(7 * B + C + N + H + W)
    x = torch.randn(B, C, H, W, generator=g).to(DT[dt])
    w = (torch.randn(N, C, 3, 3, generator=g) * 0.03).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    taps = qdm.ops.conv3x3_weight_taps(w.to(DEV))
    assert qdm.ops.conv3x3_stride2_ok(H, W)
    ref = ref_conv3x3(x, w, b, 2)
    for xin in (x.to(DEV), x.to(DEV).contiguous(memory_format=torch.channels_last)):
        y = qdm.ops.conv3x3_f16(xin, taps, b.to(DEV), stride=2)
        assert y.shape == (B, N, H // 2, W // 2) and y.dtype == DT[dt]
        assert y.is_contiguous(memory_format=torch.channels_last) or y.shape[2] * y.shape[3] == 1
        assert max_rel_err(y, ref) <= TOL
    assert max_rel_err(qdm.ops.conv3x3_f16(x.to(DEV), taps, None, stride=2), ref_conv3x3(x, w, None, 2)) <= TOL
    group = 64
    oq, oz, os_, dq = O.awq_from_linear(taps.cpu(), group, 4)
    qweight, qzeros, scales = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV)
    ref4 = ref_conv3x3(x, dq.reshape(N, 3, 3, C).permute(0, 3, 1, 2), b, 2)
    y4 = qdm.ops.conv3x3_w4a16(x.to(DEV), qweight, qzeros, scales, group, b.to(DEV), stride=2)
    assert y4.shape == (B, N, H // 2, W // 2) and max_rel_err(y4, ref4) <= TOL


def test_conv3x3_stride2_modules_and_bad_inputs(qdm):
    import importlib
    L = importlib.import_module(qdm.__name__ + ".linear")
    fq = importlib.import_module(qdm.__name__ + ".fake_quant")
    g = torch.Generator().manual_seed(5)
    conv = torch.nn.Conv2d(128, 192, 3, stride=2, padding=1).half().to(DEV)
    x = torch.randn(2, 128, 16, 16, generator=g).half().to(DEV)
    assert L.is_conv3x3_gemm(conv)
    q = L.QConv3x3.from_conv(conv, 4, L.conv_group(9 * 128, 128))
    assert q.stride == (2, 2)
    ref = torch.nn.functional.conv2d(x.float(), q.dequantize().float(), conv.bias.float(), 2, 1)
    y = q(x)
    assert y.shape == (2, 192, 8, 8) and max_rel_err(y.cpu(), ref.cpu()) <= TOL
    # a grid the tensor map cannot tile (odd, or an output width that does not divide 128): cuDNN on the dequantised weight
    for hw in ((15, 16), (16, 24)):
        xo = torch.randn(2, 128, *hw, generator=g).half().to(DEV)
        assert not qdm.ops.conv3x3_stride2_ok(*hw)
        refo = torch.nn.functional.conv2d(xo.float(), q.dequantize().float(), conv.bias.float(), 2, 1)
        assert max_rel_err(q(xo).cpu(), refo.cpu()) <= TOL
        with pytest.raises(ValueError, match="stride-2"):
            qdm.ops.conv3x3_f16(xo, qdm.ops.conv3x3_weight_taps(conv.weight.data), None, stride=2)
    with pytest.raises(ValueError, match="padded-grid"):
        qdm.ops.conv3x3_f16(x, qdm.ops.conv3x3_weight_taps(conv.weight.data), None, padded=True, stride=2)
    # fake-quant module: implicit GEMM when switched on, same result as its cuDNN path
    m = fq.WxAxConv2d.from_float(conv, weight_quant="per_channel", n_bits_W=8)
    default = fq.WxAxConv2d.conv3x3_gemm
    try:
        fq.WxAxConv2d.conv3x3_gemm = True
        assert m._conv3x3_gemm(x) and not m._conv3x3_gemm(torch.randn(2, 128, 15, 16, device=DEV).half())
        y_gemm = m(x)
        fq.WxAxConv2d.conv3x3_gemm = False
        y_dnn = m(x)
    finally:
        fq.WxAxConv2d.conv3x3_gemm = default
    assert y_gemm.shape == y_dnn.shape == (2, 192, 8, 8) and max_rel_err(y_gemm.cpu(), y_dnn.cpu()) <= TOL


def test_conv3x3_modules_and_bad_inputs(qdm):
    import importlib
    fq = importlib.import_module("quantization---diffusion-models_b200.fake_quant")
    L = importlib.import_module("quantization---diffusion-models_b200.linear")
    g = torch.Generator().manual_seed(11)
    conv = torch.nn.Conv2d(128, 192, 3, padding=1)
    conv.weight.data = (torch.randn(192, 128, 3, 3, generator=g) * 0.03).half()
    conv.bias.data = torch.randn(192, generator=g).half()
    x = torch.randn(2, 128, 12, 12, generator=g).half()
    w0, b0 = conv.weight.data.clone(), conv.bias.data.clone()
    conv = conv.to(DEV)
    # WxAxConv2d: same fake-quant weight, implicit GEMM vs cuDNN (class switch)
    m = fq.WxAxConv2d.from_float(conv, weight_quant="per_tensor", n_bits_W=8)
    assert m._conv3x3_gemm(x.to(DEV))
    default = fq.WxAxConv2d.conv3x3_gemm          # off unless QDM_CONV_GEMM=1 (cuDNN, as the reference)
    try:
        fq.WxAxConv2d.conv3x3_gemm = True
        qdm.ops.launch_count(reset=True)
        y_gemm = m(x.to(DEV))
        assert qdm.ops.launch_count() == 1         # the implicit GEMM really ran
        fq.WxAxConv2d.conv3x3_gemm = False
        y_cudnn = m(x.to(DEV))
        assert qdm.ops.launch_count() == 1
    finally:
        fq.WxAxConv2d.conv3x3_gemm = default
    assert max_rel_err(y_gemm, y_cudnn.cpu()) <= 2e-3
    assert max_rel_err(y_gemm, O.conv2d_fake(x, m.weight.cpu(), b0, 1, 1)) <= TOL
    # QConv3x3: packed weights, dequantize() returns the conv layout of the oracle's RTN
    q = L.QConv3x3.from_conv(conv, 4, L.conv_group(9 * 128, 128))
    assert q.group_size == 128
    want = O.rtn_group(w0.permute(0, 2, 3, 1).reshape(192, -1), 128, True, 4)[0].reshape(192, 3, 3, 128).permute(0, 3, 1, 2)
    assert torch.equal(q.dequantize().cpu(), want)
    assert max_rel_err(q(x.to(DEV)), O.conv2d_fake(x, want, b0, 1, 1)) <= TOL
    assert L.conv_group(9 * 320, 128) == 64 and L.conv_group(9 * 640, 128) == 128 and L.conv_group(9 * 64, 0) == 64
    with pytest.raises(ValueError):
        L.QConv3x3.from_conv(torch.nn.Conv2d(128, 192, 3, padding=1, stride=3).to(DEV).half(), 4, 64)
    with pytest.raises(ValueError):
        qdm.ops.conv3x3_f16(torch.zeros(1, 100, 4, 4, dtype=torch.float16, device=DEV),
                            torch.zeros(8, 900, dtype=torch.float16, device=DEV))     # C % 64 != 0
    with pytest.raises(ValueError):
        qdm.ops.conv3x3_f16(x.to(DEV), torch.zeros(8, 64, dtype=torch.float16, device=DEV))


SKINNY_SHAPES = [(8, 1280, 2816, 128), (1, 2432, 256, 128), (5, 72, 192, 64), (32, 640, 1280, 128), (17, 320, 320, 64),
                 (9, 8, 64, 64), (24, 1288, 1280, 128), (16, 320, 1280, 256), (16, 1280, 1280, 128), (1, 4864, 2432, 128)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K,group", SKINNY_SHAPES)
def test_gemm_w4a16_skinny_forced(qdm, dt, M, N, K, group):
    """The cluster-split-K mma.sync kernel (qdm_gemm_skinny.cu) on every shape class it accepts -- ragged N (not a
    multiple of 64 / 256), one k16 step per warp, 1..4 m-tiles, group 64 / 128 / 256 -- forced with
    set_w4_disable(W4_NO_SMALLM) (the dispatcher otherwise keeps the one-word-column kernel below ~2 M weights; the
    QDM_W4_NO_* environment switches are read once per process and cannot force anything here).  Deterministic call to call."""
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DT[dt])
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt])
    oq, oz, os_, dq = O.awq_from_linear(w, group, 4)
    qweight, qzeros, scales = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV)
    try:
        qdm.ops.set_w4_disable(qdm.ops.W4_NO_SMALLM)
        y1 = qdm.ops.gemm_w4a16(x.to(DEV), qweight, qzeros, scales, group, b.to(DEV))
        assert qdm.ops.gemm_last_variant()[0] == "skinny", qdm.ops.gemm_last_variant()
        y2 = qdm.ops.gemm_w4a16(x.to(DEV), qweight, qzeros, scales, group, b.to(DEV))
    finally:
        qdm.ops.set_w4_disable(0)
    assert y1.shape == (M, N) and y1.dtype == DT[dt] and torch.equal(y1, y2)
    assert max_rel_err(y1, ref_linear(x, dq, b)) <= TOL
    y_kn = qdm.ops.gemm_f16_kn(x.to(DEV), qdm.ops.dequant_awq(qweight, qzeros, scales, group), b.to(DEV))
    assert max_rel_err(y1, y_kn.cpu()) <= (2e-3 if dt == "f16" else 8e-3)


def test_gemm_w4a16_skinny_vs_other_kernels(qdm):
    """The three W4A16 kernels an M <= 32 problem can take agree up to accumulation order on the same packed weights:
    skinny (forced), the one-word-column small-M kernel (where it fits) and the tcgen05 kernel."""
    g = torch.Generator().manual_seed(21)
    for M, N, K, group in [(16, 1280, 1280, 128), (2, 14592, 2432, 128)]:
        x = torch.randn(M, K, generator=g).half().to(DEV)
        w = (torch.randn(N, K, generator=g) * 0.05).half().to(DEV)
        b = torch.randn(N, generator=g).half().to(DEV)
        qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w, group, want_dq=True)
        try:
            qdm.ops.set_w4_disable(qdm.ops.W4_NO_SMALLM)
            y_sk = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)          # skinny
            assert qdm.ops.gemm_last_variant()[0] == "skinny"
            qdm.ops.set_w4_disable(qdm.ops.W4_NO_SMALLM | qdm.ops.W4_NO_SKINNY)
            y_tc = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)          # tcgen05
            assert qdm.ops.gemm_last_variant()[0] in ("single", "pair", "streamk")
            qdm.ops.set_w4_disable(qdm.ops.W4_NO_SKINNY)
            y_sm = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b)          # small-M where it fits, else tcgen05
            assert qdm.ops.gemm_last_variant()[0] == ("smallm" if N * K <= (13 << 19) else qdm.ops.gemm_last_variant()[0])
        finally:
            qdm.ops.set_w4_disable(0)
        assert max_rel_err(y_sk, ref_linear(x.cpu(), dq.cpu(), b.cpu())) <= TOL
        assert max_rel_err(y_sk, y_tc.cpu()) <= 2e-3 and max_rel_err(y_sk, y_sm.cpu()) <= 2e-3
