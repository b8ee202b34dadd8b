"""GPU parity of kernels (c) W4A16 and (d) W8A8 on EVERY Linear shape of the BASELINE.json configs and on the
repacked-weight ("RP") kernel, through the C ABI.

  * every distinct (M, N, K) of the SD1.5 UNet (batch 8 + CFG), the SDXL UNet (batch 4 + CFG) and the SD3.5-L MMDiT
    (batch 1) Linear inventories (shapes.py = SURVEY.md Appendix A) and the corners of the config-5 sweep, fp16 and bf16;
  * reference: the reference's own formulation on the same inputs, F.linear(x, W_fakequant, bias)
    (quantize/fake_quant.py:223), evaluated in fp32 (no TF32) -- tolerance max |y - ref| / max |ref| <= 1e-2 (north star);
  * the kernel variant the dispatcher picked is asserted (qdm_gemm_last_variant), because B-stationary, wide-tile and
    small-M dispatch is shape dependent.
"""
import importlib

import pytest
import torch

from _util import DT, max_rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2

shapes = importlib.import_module("quantization---diffusion-models_b200.shapes")


def _distinct(layers):
    return sorted({(m, n, k) for _, m, n, k, _ in layers}, reverse=True)


SD15 = _distinct(shapes.sd15_unet_linears(batch=8, cfg=True))
SDXL = _distinct(shapes.sdxl_unet_linears(batch=4, cfg=True))
SD35 = _distinct(shapes.sd35_mmdit_linears(batch=1))
SWEEP = [(4096, 1536, 1536), (4096, 6144, 6144), (65536, 1536, 1536), (65536, 6144, 6144),
         (4096, 6144, 1536), (4096, 1536, 6144), (16384, 8192, 2048), (16384, 2048, 8192)]
ALL = [("sd15",) + s for s in SD15] + [("sdxl",) + s for s in SDXL if s not in SD15] + \
      [("sd35",) + s for s in SD35] + [("sweep",) + s for s in SWEEP]


def ref_linear_gpu(x, w_nk, bias):
    """fp32 reference on the GPU in row chunks (M up to 65536 x N up to 10240 would be 2.7 GB in one piece)"""
    assert not torch.backends.cuda.matmul.allow_tf32
    wt = w_nk.float().t().contiguous()
    outs = []
    for i in range(0, x.shape[0], 8192):
        y = x[i:i + 8192].float() @ wt
        if bias is not None:
            y = y + bias.float()
        outs.append(y)
    return torch.cat(outs)


def rel_err_gpu(y, ref):
    return ((y.float() - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def expected_variants(M, N, K):
    if M <= 32:
        return {"skinny", "smallm"}
    if M <= 128:
        return {"single", "ts"}
    return {"pair", "bstat", "streamk", "ts"}


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("model,M,N,K", ALL)
def test_w4a16_baseline_shape(qdm, model, M, N, K, dt):
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + 7 * K)
    grp = shapes.group_for(K)
    x = torch.randn(M, K, generator=g, device=DEV, dtype=DT[dt])
    w = (torch.randn(N, K, generator=g, device=DEV) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g, device=DEV).to(DT[dt])
    lin = torch.nn.Linear(K, N, bias=True, device=DEV, dtype=DT[dt])
    lin.weight.data, lin.bias.data = w, b
    linear = importlib.import_module("quantization---diffusion-models_b200.linear")
    mod = linear.WQLinear_GEMM.from_linear(lin, 4, grp)
    y = mod(x)
    variant, tile = qdm.ops.gemm_last_variant()
    assert variant in expected_variants(M, N, K), (variant, tile)
    assert y.shape == (M, N) and y.dtype == DT[dt]
    dq = mod.dequantize()                                   # [N, K] fake-quant weight (bit-exact with the oracle, test_gpu_quant)
    ref = ref_linear_gpu(x, dq, b)
    err = rel_err_gpu(y, ref)
    assert err <= TOL, (variant, tile, err)
    assert torch.isfinite(y).all()
    # the module's dispatch: the tensor-memory-A kernel everywhere except small K with many tiles (B-stationary kernel) and
    # N = 320 with many token tiles (AWQ-tensor kernel)
    if (M, N, K) in ((4096, 1280, 1280), (8192, 1280, 1280), (4096, 2432, 2432), (1232, 1280, 768), (333, 2432, 2432), (4096, 10240, 1280),
                     (4096, 9728, 2432), (16384, 5120, 640)):
        assert variant == "ts", (variant, tile)
    if (M, N, K) in ((4096, 1280, 5120),):                                            # one wave less on a long main loop: wide tile
        assert variant == "ts" and tile > 192, (variant, tile)
    if (M, N, K) in ((4096, 10240, 1280), (16384, 5120, 640), (8192, 1280, 1280), (4096, 9728, 2432)):   # overlapped epilogue
        assert variant == "ts" and tile <= 192, (variant, tile)
    if (M, N, K) in ((16384, 640, 2560), (32768, 640, 2560), (16384, 640, 640)):      # N = 640 with many token tiles
        assert variant == "pair", (variant, tile)
    if (M, N, K) in ((65536, 320, 320), (65536, 2560, 320), (65536, 320, 1280)):      # K = 320 / N = 320 with many token tiles
        assert variant in ("pair", "bstat"), (variant, tile)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("model,M,N,K", [a for a in ALL if a[3] % 16 == 0])
def test_w8a8_baseline_shape(qdm, model, M, N, K, dt):
    import oracle.qdm_oracle as O
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + 7 * K + 1)
    x = torch.randn(M, K, generator=g, device=DEV)
    x[:, 3] *= 20                                           # an outlier channel, what SmoothQuant is for
    x = x.to(DT[dt])
    w = (torch.randn(N, K, generator=g, device=DEV) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g, device=DEV).to(DT[dt])
    xq, sx = qdm.ops.actquant_token_i8(x)
    _, wq, sw, _ = qdm.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
    y = qdm.ops.gemm_w8a8(xq, sx, wq, sw.float(), b, out_dtype=DT[dt])
    # exact integer reference of the same codes, fp32 scales (products of int8 sums up to K * 127^2 < 2^31 are exact in
    # fp64 and, row-chunked, cheap on the GPU)
    wqt = wq.double().t().contiguous()
    worst = 0.0
    for i in range(0, M, 4096):
        yi = (xq[i:i + 4096].double() @ wqt) * sx[i:i + 4096].double()[:, None] * sw.double()[None, :] + b.double()
        worst = max(worst, ((y[i:i + 4096].double() - yi).abs().max() / yi.abs().max().clamp_min(1e-12)).item())
    assert worst <= 4e-3, worst                              # one rounding to the output dtype (bf16: 2^-8)
    # the reference's fake-quant formulation (fake_quant.py:86-93,109-118,223) on a row sample the CPU oracle finishes fast.
    # fp16 only: that is the dtype the reference's path computes in (WxAxLinear keeps fp16 weights, fake_quant.py:179); in
    # bf16 the formulation itself rounds every q * s product to 8 bits, which is its noise, not the kernel's (the exact
    # integer check above covers bf16)
    if dt == "f16":
        rows = torch.linspace(0, M - 1, steps=min(M, 96)).long()
        ref = O.linear_w8a8_fake(x[rows.to(DEV)].cpu(), w.cpu(), b.cpu())
        assert max_rel_err(y[rows.to(DEV)], ref) <= TOL


RP_CASES = [  # (M, N, K, group): ragged M / N, N % 32 != 0, odd k-block counts, group 64 / 128 / 256
    (300, 320, 320, 64), (513, 72, 192, 64), (777, 96, 448, 64), (1024, 2432, 2432, 128), (4096, 1280, 1280, 128),
    (1232, 1280, 768, 128), (2048, 64, 2432, 128), (4096, 2560, 320, 64), (333, 9728, 2432, 128), (616, 1280, 2048, 256),
    (129, 520, 256, 128), (8192, 1280, 1280, 128), (1000, 40, 192, 64),
]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K,group", RP_CASES)
def test_w4a16_repacked_kernel(qdm, M, N, K, group, dt):
    """The repacked-weight kernel, both forms (one / two sub-tiles per CTA pair) and several pinned sub-tile widths,
    against the fp32 reference and against the kernel that reads the AWQ tensors directly (same dequantised values,
    so only the fp32 summation order may differ: <= one output ulp at the largest magnitude)."""
    import oracle.qdm_oracle as O
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DT[dt]).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt]).to(DEV)
    if N % 64 == 0:
        qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w.to(DEV), group, want_dq=True)
    else:
        oq, oz, os_, dq = O.awq_from_linear(w, group, 4)
        qweight, qzeros, scales, dq = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV), dq.to(DEV)
    blob = qdm.ops.w4a16_repack(qweight, qzeros, scales, group)
    ref = ref_linear_gpu(x, dq, b)
    ulp = 2e-3 if dt == "f16" else 8e-3
    try:
        qdm.ops.set_gemm_mode(64)                            # the AWQ-tensor kernels, blob ignored
        y_awq = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, blob)
        assert qdm.ops.gemm_last_variant()[0] not in ("rp1", "rp2")
        for mode, name in ((16, "rp1"), (32, "rp2")):
            for width in (0, 32, 96, 160, 256):
                if name == "rp2" and width and width >= N:
                    continue
                qdm.ops.set_gemm_mode(mode | (width << 8))
                y1 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, blob)
                variant, tile = qdm.ops.gemm_last_variant()
                assert variant == name and (width == 0 or tile == width * (2 if name == "rp2" else 1)), (variant, tile)
                y2 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, blob)
                assert torch.equal(y1, y2), (name, width)    # deterministic
                assert rel_err_gpu(y1, ref) <= TOL, (name, width)
                assert rel_err_gpu(y1, y_awq.float()) <= ulp, (name, width)
    finally:
        qdm.ops.set_gemm_mode(0)


def test_w4a16_repack_layout(qdm):
    """The blob is a pure re-arrangement: block (kg, nb) holds qweight[128 kg + k, 2 nb + w] at k * 8 + w * 4, the scales of
    the two 64-row halves at 1024 and the packed zero points at 1088 (include/qdm.h: qdm_w4a16_repack)."""
    g = torch.Generator().manual_seed(5)
    for N, K, group in [(72, 192, 64), (320, 320, 64), (256, 512, 128), (520, 1024, 256)]:
        qweight = torch.randint(-2 ** 31, 2 ** 31 - 1, (K, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
        qzeros = torch.randint(-2 ** 31, 2 ** 31 - 1, (K // group, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
        scales = torch.randn(K // group, N, generator=g).half()
        blob = qdm.ops.w4a16_repack(qweight.to(DEV), qzeros.to(DEV), scales.to(DEV), group).cpu()
        nb, kg = (N + 15) // 16, (K + 127) // 128
        assert blob.numel() == nb * kg * 1104
        blk = blob.view(kg, nb, 1104)
        words = blk[:, :, :1024].contiguous().view(torch.int32).view(kg, nb, 128, 2)
        sc = blk[:, :, 1024:1088].contiguous().view(torch.float16).view(kg, nb, 2, 16)
        zw = blk[:, :, 1088:1104].contiguous().view(torch.int32).view(kg, nb, 2, 2)
        qw_pad = torch.zeros(kg * 128, nb * 2, dtype=torch.int32)
        qw_pad[:K, :N // 8] = qweight
        assert torch.equal(words, qw_pad.view(kg, 128, nb, 2).permute(0, 2, 1, 3))
        for a in range(kg):
            for h in range(2):
                kk = a * 128 + 64 * h
                if kk >= K:
                    assert not sc[a, :, h].any() and not zw[a, :, h].any()
                    continue
                s_pad = torch.zeros(nb * 16, dtype=torch.float16)
                s_pad[:N] = scales[kk // group]
                z_pad = torch.zeros(nb * 2, dtype=torch.int32)
                z_pad[:N // 8] = qzeros[kk // group]
                assert torch.equal(sc[a, :, h].reshape(-1), s_pad)
                assert torch.equal(zw[a, :, h].reshape(-1), z_pad)


TS_CASES = [(256, 256, 128, 128), (300, 320, 320, 64), (777, 520, 448, 64), (4096, 1280, 1280, 128), (1232, 1280, 768, 128),
            (333, 2432, 2432, 128), (4096, 64, 2432, 128), (8192, 640, 2560, 128), (129, 72, 256, 128), (200, 8, 128, 128),
            (64, 1280, 320, 64), (4096, 2560, 320, 64), (616, 1280, 2048, 256), (1000, 40, 192, 64)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,N,K,group", TS_CASES)
def test_w4a16_tmem_a_kernel(qdm, M, N, K, group, dt):
    """The kernel that feeds the dequantised weights to tcgen05.mma as its A operand from tensor memory (output channels on
    the TMEM lanes, tokens along N), at every token-tile width, against the fp32 reference and against the kernels that
    read the AWQ tensors (same dequantised values: only the fp32 summation order may differ, <= one output ulp at the
    largest magnitude); ragged M / N / K-tail shapes included; twice for determinism."""
    import oracle.qdm_oracle as O
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(DT[dt]).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DT[dt])
    b = torch.randn(N, generator=g).to(DT[dt]).to(DEV)
    if N % 64 == 0:
        qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w.to(DEV), group, want_dq=True)
    else:
        oq, oz, os_, dq = O.awq_from_linear(w, group, 4)
        qweight, qzeros, scales, dq = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV), dq.to(DEV)
    bts = qdm.ops.w4a16_repack_ts(qweight, qzeros, scales, group)
    ref = ref_linear_gpu(x, dq, b)
    ulp = 2e-3 if dt == "f16" else 8e-3
    try:
        qdm.ops.set_gemm_mode(64)                            # the AWQ-tensor kernels
        y_awq = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, None, bts)
        assert qdm.ops.gemm_last_variant()[0] != "ts"
        for width in (0, 32, 64, 96, 128, 160, 192, 256, 320, 384):   # > 192: one wide tile of two sub-tiles
            qdm.ops.set_gemm_mode(128 | (width << 8))
            y1 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, None, bts)
            variant, tile = qdm.ops.gemm_last_variant()
            assert variant == "ts" and (width == 0 or tile == width), (variant, tile)
            y2 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, None, None, bts)      # also without bias
            y3 = qdm.ops.gemm_w4a16(x, qweight, qzeros, scales, group, b, None, bts)
            assert torch.equal(y1, y3), width                # deterministic
            assert rel_err_gpu(y1, ref) <= TOL, width
            assert rel_err_gpu(y1, y_awq.float()) <= ulp, width
            assert rel_err_gpu(y2, ref - b.float()) <= TOL, width
    finally:
        qdm.ops.set_gemm_mode(0)


@pytest.mark.parametrize("case", [(300, 512, 256, 0, True), (1232, 1280, 768, 0, True), (4096, 320, 320, 0, True), (333, 2432, 2432, 0, True),
                                  (512, 640, 640, 2, False), (65536, 320, 320, 0, False), (4096, 1280, 2560, 8, False), (100, 256, 128, 1, False),
                                  (16, 1280, 320, 0, True), (1, 4864, 2432, 0, True), (129, 72, 256, 0, True)])
def test_w4a16_writes_only_its_output(qdm, case):
    """Own bounds check (compute-sanitizer is closed on this pool, profiles/sanitizer_r02.txt): every kernel family writes
    into a caller-owned [M, N] view in the middle of a sentinel-filled buffer; the guard rows before and after must survive
    and the result must still match the fp32 reference (a column overrun would land in the next row)."""
    M, N, K, mode, ts = case
    g = torch.Generator(device=DEV).manual_seed(11 * M + N + K)
    grp = shapes.group_for(K)
    x = torch.randn(M, K, generator=g, device=DEV, dtype=torch.float16)
    w = (torch.randn(N, K, generator=g, device=DEV) * 0.05).half()
    b = torch.randn(N, generator=g, device=DEV).half()
    if N % 64 == 0:
        qw, qz, sc, dq = qdm.ops.quant_pack_awq(w, grp, want_dq=True)
    else:
        import oracle.qdm_oracle as O
        oq, oz, os_, dq = O.awq_from_linear(w.cpu(), grp, 4)
        qw, qz, sc, dq = torch.from_numpy(oq).to(DEV), torch.from_numpy(oz).to(DEV), os_.to(DEV), dq.to(DEV)
    bts = qdm.ops.w4a16_repack_ts(qw, qz, sc, grp) if ts else None
    guard = 8
    big = torch.full((M + 2 * guard, N), 12345.0, dtype=torch.float16, device=DEV)
    out = big[guard:guard + M]
    qdm.ops.set_gemm_mode(mode)
    try:
        y = qdm.ops.gemm_w4a16(x, qw, qz, sc, grp, b, None, bts, out=out)
    finally:
        qdm.ops.set_gemm_mode(0)
    assert y.data_ptr() == out.data_ptr()
    assert bool((big[:guard] == 12345.0).all()) and bool((big[guard + M:] == 12345.0).all()), qdm.ops.gemm_last_variant()
    assert rel_err_gpu(y, ref_linear_gpu(x, dq, b)) <= TOL, qdm.ops.gemm_last_variant()


def test_w4a16_ts_panel_tile_order():
    """The TS kernel's tile order in panels of `group_m` token blocks (qdm_gemm_w4ts.cu, ts_nblk / ts_mblk): forced panel
    heights through the read-once QDM_W4_TS_GM knob in a fresh process -- 1 (channel blocks fastest), 3 (a ragged last
    panel: 5 token blocks of 192 = 3 + 2) and -1 (no panels) must all give the fp32 reference's result, and identical bits."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import importlib, sys, torch
sys.path.insert(0, %r)
q = importlib.import_module("quantization---diffusion-models_b200")
g = torch.Generator(device="cuda:0").manual_seed(3)
out = []
for (M, N, K, tile) in ((900, 712, 384, 192), (1500, 520, 256, 128), (1100, 264, 512, 384)):
    x = torch.randn(M, K, generator=g, device="cuda:0", dtype=torch.float16)
    w = (torch.randn(N, K, generator=g, device="cuda:0") * 0.05).half()
    import oracle.qdm_oracle as O
    oq, oz, os_, dq = O.awq_from_linear(w.cpu(), 128, 4)
    qw, qz, sc = torch.from_numpy(oq).cuda(), torch.from_numpy(oz).cuda(), os_.cuda()
    bts = q.ops.w4a16_repack_ts(qw, qz, sc, 128)
    q.ops.set_gemm_mode(128 | (tile << 8))
    y = q.ops.gemm_w4a16(x, qw, qz, sc, 128, None, None, bts)
    assert tuple(q.ops.gemm_last_variant()) == ("ts", tile), q.ops.gemm_last_variant()
    ref = x.float() @ dq.cuda().float().t()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    assert err <= 1e-2, err
    out.append(int(y.view(torch.int16).long().sum().item()))
print("SUMS", out)
''' % root
    sums = {}
    for v in ("0", "1", "3", "-1"):
        env = dict(os.environ, QDM_W4_TS_GM=v)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        sums[v] = [l for l in r.stdout.splitlines() if l.startswith("SUMS")][-1]
    assert sums["1"] == sums["3"] == sums["-1"] == sums["0"]          # the tile order does not change a single output bit
