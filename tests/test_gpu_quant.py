"""GPU parity for kernels (a) reductions and (b) quantise/pack: CUDA path (through the C ABI) vs the CPU
oracle and vs the reference-generated golden fixtures.  Bit-exact (torch.equal) everywhere except the
fp32-accumulated sums, whose tolerance is stated at the assert."""
import numpy as np
import pytest
import torch

from _util import DT, Golden, assert_bit_equal

import oracle.qdm_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rand_w(shape, dtype, seed, scale=0.05):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(shape, generator=g) * scale
    if len(shape) == 2 and shape[0] > 8:
        w[1, : shape[1] // 2] *= 40.0
        w[2] = 0.0
        w[3] = 0.0123
        w[5] = w[5].abs()
        w[6] = -w[6].abs()
    return w.to(dtype)


# ------------------------------------------------------------------ golden fixtures
def test_golden_pseudo_quantize_tensor(qdm):
    g = Golden("pseudo_quantize_tensor.npz")
    for tag, dt, group, zp, bits in g.cases():
        w = g.get(tag + "_w").to(DEV)
        dq, codes, s, z = qdm.ops.quant_group(w, int(group), int(bits), zero_point=bool(int(zp)), want_codes=True)
        assert_bit_equal(dq, g.get(tag + "_dq"), f"{tag} dq")
        assert_bit_equal(s, g.get(tag + "_s"), f"{tag} scales")
        if int(zp):
            assert_bit_equal(z, g.get(tag + "_z"), f"{tag} zeros")
        oc = O.rtn_group(g.get(tag + "_w"), int(group), bool(int(zp)), int(bits))[3]
        assert torch.equal(codes.cpu().to(torch.int32), oc.to(torch.int32)), f"{tag} codes"


def test_golden_fake_quant(qdm):
    from importlib import import_module
    fq = import_module("quantization---diffusion-models_b200.fake_quant")
    g = Golden("fake_quant.npz")
    for tag, kind, dt, group, bits in g.cases():
        w, want = g.get(tag + "_w").to(DEV), g.get(tag + "_dq")
        if kind == "group":
            got = fq.quantize_weight_absmax(w.clone(), int(bits), int(group), codeBookQuantInd=False)
        elif kind == "channel":
            got = fq.quantize_weight_per_channel_absmax(w, int(bits))
        elif kind == "token":
            got = fq.quantize_activation_per_token_absmax(w, int(bits))
        elif kind == "tensor":
            got = fq.quantize_weight_per_tensor_absmax(w, int(bits))
        elif kind == "nchw":
            got = fq.quantize_activation_per_channel_absmax(w, int(bits))
        assert_bit_equal(got, want, f"{tag} {kind}")


def test_golden_activation_quantisers_round2(qdm):
    """quantize_activation_per_channel_group_absmax (fake_quant.py:134-153: per (n, c, patch), incl. the `group_size -= 2`
    fallback) and the 16-bit per-token / per-tensor / per-(n, c) forms (the reference's default a_bit = 16: q_max = 32767)
    against the reference-generated fixture, bit for bit."""
    import importlib
    fq = importlib.import_module("quantization---diffusion-models_b200.fake_quant")
    g = Golden("act_quant.npz")
    for tag, kind, dt, gs, bits in g.cases():
        x, want = g.get(tag + "_x").to(DEV), g.get(tag + "_y")
        got = {"patch": lambda: fq.quantize_activation_per_channel_group_absmax(x, group_size=int(gs), n_bits=int(bits)),
               "token": lambda: fq.quantize_activation_per_token_absmax(x, int(bits)),
               "tensor": lambda: fq.quantize_activation_per_tensor_absmax(x, int(bits)),
               "nchw": lambda: fq.quantize_activation_per_channel_absmax(x, int(bits))}[kind]()
        assert_bit_equal(got, want, f"{tag} {kind} {dt}")


def test_golden_awq_layout(qdm):
    g = Golden("awq_layout.npz")
    codes_kn = g.get("codes").to(torch.int8)
    qweight = qdm.ops.pack_awq(codes_kn.t().contiguous().to(DEV))
    assert torch.equal(qweight.cpu(), g.get("qweight"))
    assert torch.equal(qdm.ops.unpack_awq(g.get("qweight").to(DEV)).cpu(), codes_kn)
    deq = qdm.ops.dequant_awq(g.get("qweight").to(DEV), g.get("qzeros").to(DEV), g.get("scales").to(DEV), int(g.get("group")))
    assert_bit_equal(deq, g.get("deq"), "dequantize_gemm")


# ------------------------------------------------------------------ (b) vs oracle, wider sweep
@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("group,zp,bits", [(128, True, 4), (64, True, 4), (128, False, 4), (128, True, 8),
                                           (32, True, 3), (256, False, 8), (96, True, 4), (0, True, 4)])
def test_quant_group_vs_oracle(qdm, dt, group, zp, bits):
    K = 768 if group != 256 else 1024
    w = rand_w((160, K), DT[dt], 11 + bits + group)
    dq, codes, s, z = qdm.ops.quant_group(w.to(DEV), group, bits, zero_point=zp, want_codes=True)
    odq, os_, oz, oc = O.rtn_group(w, group, zp, bits)
    assert_bit_equal(dq, odq, "dq")
    assert_bit_equal(s, os_, "scales")
    if zp:
        assert_bit_equal(z, oz, "zeros")
    assert torch.equal(codes.cpu().to(torch.int32), oc.to(torch.int32))


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_group_search_fusions(qdm, dt):
    """pre_mul / clip / post_div are the fused W*s, clamp and /s of quantizer.py:727-730,845."""
    w = rand_w((96, 512), DT[dt], 5)
    g = torch.Generator().manual_seed(6)
    s = (torch.rand(512, generator=g) * 2 + 0.25).to(DT[dt])
    clip = (torch.rand(96 * 4, generator=g) * 0.1 + 0.02).to(DT[dt])
    dq, _, sc, z = qdm.ops.quant_group(w.to(DEV), 128, 4, True, pre_mul=s.to(DEV), post_div=s.to(DEV))
    ref = O.rtn_group(w * s.view(1, -1), 128, True, 4)
    assert_bit_equal(dq, ref[0] / s.view(1, -1), "Q(W*s)/s")
    assert_bit_equal(sc, ref[1], "scales")
    dq2 = qdm.ops.quant_group(w.to(DEV), 128, 4, True, clip_max=clip.to(DEV))[0]
    wc = O.apply_clip(w, clip.view(96, 4, 1))
    assert_bit_equal(dq2, O.rtn_group(wc, 128, True, 4)[0], "Q(clamp(W))")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("shape", [(320, 1280), (64, 32, 3, 3), (7, 5, 1, 1), (33, 100), (4, 77, 320), (5, 3, 3, 3), (16, 4, 5, 5), (9, 6)])
def test_rowwise_vs_oracle(qdm, dt, shape):
    x = rand_w(shape, DT[dt], 3, scale=1.0)
    dq, codes, s, _ = qdm.ops.quant_rowwise(x.to(DEV), 8, want_codes=True, want_scales=True)
    odq, oc, os_ = O.rtn_rows(x, 8)
    assert_bit_equal(dq, odq, "dq")
    assert_bit_equal(s, os_.reshape(-1), "scales")
    assert torch.equal(codes.cpu().to(torch.int32), oc.clamp(-128, 127).to(torch.int32))
    if dt == "f16":
        assert oc.abs().max() <= 127  # fp16 never reaches +-128 (DESIGN.md, 8-bit caveat)
    # without integer codes the tiny-row shapes (conv kw = 3 / 1) take the vectorised kernel: same bits
    dq2, _, s2, _ = qdm.ops.quant_rowwise(x.to(DEV), 8, want_scales=True)
    assert torch.equal(dq2.cpu().view(torch.int16), odq.view(torch.int16)), "dq bits (no-codes path)"
    assert_bit_equal(s2, os_.reshape(-1), "scales (no-codes path)")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_tensor_and_absmax(qdm, dt):
    x = rand_w((37, 129), DT[dt], 4, scale=2.0)
    dq, codes, s = qdm.ops.quant_tensor(x.to(DEV), 8, want_codes=True)
    odq, oc, os_ = O.rtn_tensor(x, 8)
    assert_bit_equal(dq, odq, "dq")
    assert_bit_equal(s, os_, "scale")
    assert_bit_equal(qdm.ops.absmax(x.to(DEV)), x.abs().max(), "absmax")
    assert_bit_equal(qdm.ops.rowabsmax(x.to(DEV)), x.abs().max(dim=-1)[0], "rowabsmax")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_actquant_token_i8(qdm, dt):
    x = rand_w((300, 640), DT[dt], 8, scale=1.5)
    xq, sx = qdm.ops.actquant_token_i8(x.to(DEV))
    _, oc, os_ = O.rtn_rows(x, 8)
    assert torch.equal(xq.cpu().to(torch.int32), oc.clamp(-128, 127).to(torch.int32))
    assert torch.equal(sx.cpu(), os_.reshape(-1).float())
    sm = (torch.rand(640) + 0.5).to(DT[dt])
    xq2, sx2 = qdm.ops.actquant_token_i8(x.to(DEV), smooth=sm.to(DEV))
    _, oc2, os2 = O.rtn_rows(x / sm, 8)
    assert torch.equal(xq2.cpu().to(torch.int32), oc2.clamp(-128, 127).to(torch.int32))
    assert torch.equal(sx2.cpu(), os2.reshape(-1).float())


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("N,K,group", [(128, 256, 128), (192, 320, 64), (64, 128, 32), (2432, 2432, 128), (72, 192, 64),
                                         (320, 320, 64), (8, 256, 256), (1288, 384, 128)])
def test_quant_pack_awq_vs_oracle(qdm, dt, N, K, group):
    w = rand_w((N, K), DT[dt], 21)
    qweight, qzeros, scales, dq = qdm.ops.quant_pack_awq(w.to(DEV), group, want_dq=True)
    oq, oz, os_, odq = O.awq_from_linear(w, group, 4)
    assert np.array_equal(qweight.cpu().numpy(), oq)
    assert np.array_equal(qzeros.cpu().numpy(), oz)
    assert_bit_equal(scales, os_, "scales_t")
    assert_bit_equal(dq, odq, "dq")
    # decode -> the fake-quant weight again (layout round trip, packing_utils.py:87-102)
    deq = qdm.ops.dequant_awq(qweight, qzeros, scales, group)
    assert_bit_equal(deq, O.awq_dequant(oq, oz, os_, group), "dequant")
    assert_bit_equal(deq.t().contiguous(), odq, "dequant == fake-quant weight")


def test_fastdiv_exhaustive(qdm):
    """The quantise kernels divide by `reciprocal + two FMAs` instead of IEEE division.  The library sweeps every
    (dividend, divisor) pair of 16-bit values inside the window where that path is taken and compares with
    __fdiv_rn after the dtype rounding: 0 differences allowed (csrc/qdm_common.cuh)."""
    import ctypes
    for dt, min_pairs in ((qdm._lib.QDM_F16, 31743 * 63488), (qdm._lib.QDM_BF16, 6 * 10 ** 8)):
        out = (ctypes.c_uint64 * 2)()
        qdm._lib.check(qdm._lib.load().qdm_selftest_fastdiv(dt, ctypes.cast(out, ctypes.c_void_p)))
        assert out[0] >= min_pairs, out[0]
        assert out[1] == 0, f"{out[1]} of {out[0]} quotients differ from IEEE division"


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_group_wide_range_and_signed_zero(qdm, dt):
    """Rows whose groups sit far from the unit range (tiny, huge, constant, one-sided) and the sign of the zeros
    torch.round produces in the symmetric modes: bit patterns, not just values, must match the oracle."""
    g = torch.Generator().manual_seed(77)
    w = torch.randn(64, 512, generator=g)
    w[0] *= 1e-7; w[1] *= 5e3 if dt == "f16" else 1e30; w[2] = 1.0; w[3] = w[3].abs() + 5.0
    w[4] *= 1e-3; w[5, ::2] = 0.0; w[6] = -1e-6; w[7] = (w[7].abs() * 1e4 + 1e4) if dt == "f16" else w[7] * 1e37
    w[8, ::3] = -0.0; w[9] = 0.0; w[10] = -0.0
    w = w.clamp(-6e4, 6e4).to(DT[dt]) if dt == "f16" else w.to(DT[dt])
    for zp, nc in ((True, False), (False, False), (False, True)):
        dq = qdm.ops.quant_group(w.to(DEV), 128, 4, zero_point=zp, no_clamp=nc)[0]
        if nc:                                        # fake_quant.py:75 returns fp16 whatever the input dtype
            dq, ref = dq.to(torch.float16), O.rtn_absmax_group(w, 4, 128)[0]
        else:
            ref = O.rtn_group(w, 128, zp, 4)[0]
        dq = dq.cpu()
        assert torch.equal(dq.isnan(), ref.isnan())
        ok = ~ref.isnan()
        a, b = dq.view(torch.int16)[ok], ref.view(torch.int16)[ok]
        assert torch.equal(a, b), f"zp={zp} nc={nc}: {(a != b).sum().item()} bit patterns differ"


# ------------------------------------------------------------------ (a) reductions
@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("rows,cols", [(4096, 2432), (77, 768), (1, 320), (1000, 20), (513, 1283)])
def test_col_reductions(qdm, dt, rows, cols):
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g)
    x[:, cols // 3] *= 50
    x = x.to(DT[dt])
    assert_bit_equal(qdm.ops.colabsmax(x.to(DEV)), O.hook_colabsmax(x), "colabsmax")
    run = qdm.ops.colabsmax(x.to(DEV))
    x2 = (x * 0.5).to(DT[dt])
    x2[0, 0] = 1000.0
    qdm.ops.colabsmax(x2.to(DEV), out=run, running=True)
    assert_bit_equal(run, torch.maximum(O.hook_colabsmax(x), O.hook_colabsmax(x2)), "running max")
    ssum = qdm.ops.colabssum(x.to(DEV)).cpu()
    ref = x.abs().double().sum(0)
    # fp32 accumulation in a fixed tree vs fp64: relative error bound 1e-5
    assert ((ssum.double() - ref).abs() / ref.clamp_min(1e-30)).max().item() < 1e-5
    # deterministic run to run
    assert torch.equal(ssum, qdm.ops.colabssum(x.to(DEV)).cpu())
    if dt != "f32":
        xm = (qdm.ops.colabssum(x.to(DEV)) / rows).to(DT[dt]).cpu()
        om = O.awq_x_mean(x)
        # x_mean rounds an fp32 mean once to 16 bits; a different fp32 summation order can flip a value
        # sitting on a rounding boundary by 1 ulp -- allow that on at most 0.1 % of channels.
        assert (xm != om).float().mean().item() <= 1e-3
        assert ((xm.float() - om.float()).abs() <= om.float().abs() * 2 ** -7).all()


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_awq_wmean(qdm, dt):
    w = rand_w((3 * 320, 768), DT[dt], 9)
    got = (qdm.ops.awq_wsum(w.to(DEV), 128) / w.shape[0]).to(DT[dt]).cpu()
    want = O.awq_w_mean([w], 128)
    # fp32 sum in a different (fixed) order, rounded once to 16 bits: a value on a rounding boundary may
    # land 1 ulp away; seen on 1 of 768 channels -> allow 0.5 %, never more than 1 ulp (next assert)
    assert (got != want).float().mean().item() <= 5e-3
    assert ((got.float() - want.float()).abs() <= want.float().abs() * 2 ** -7).all()


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_sqdiff(qdm, dt):
    g = torch.Generator().manual_seed(2)
    a = torch.randn(4096, 333, generator=g).to(DT[dt])
    b = (a.float() + 0.01 * torch.randn(4096, 333, generator=g)).to(DT[dt])
    got = qdm.ops.sqdiff_sum(a.to(DEV), b.to(DEV)).item()
    want = (a - b).float().pow(2).double().sum().item()
    assert abs(got - want) <= 1e-5 * want
    assert abs(got / a.numel() - O.mse_loss(a, b)) <= 1e-5 * O.mse_loss(a, b)


def test_empty_and_bad_inputs(qdm):
    with pytest.raises(ValueError):
        qdm.ops.quant_group(torch.zeros(0, 128, dtype=torch.float16, device=DEV), 128)
    with pytest.raises(ValueError):
        qdm.ops.quant_group(torch.zeros(4, 100, dtype=torch.float16, device=DEV), 64)
    with pytest.raises(ValueError):
        qdm.ops.colabsmax(torch.zeros(0, 8, dtype=torch.float16, device=DEV))
    with pytest.raises(RuntimeError):
        qdm.ops.colabsmax(torch.zeros(4, 8, dtype=torch.float16))


# ------------------------------------------------------------------ one-pass hook statistic
@pytest.mark.parametrize("dt", ["f16", "bf16", "f32"])
@pytest.mark.parametrize("rows,cols", [(4096, 2432), (77, 768), (1, 320), (1000, 20), (513, 1283)])
def test_colstats_one_pass(qdm, dt, rows, cols):
    """qdm_colstats: per-call max bit-exact, fp64 sum of the per-call maxima exact, |x| sum within fp32-tree error,
    running max and accumulators folded in place over several calls, any subset of the outputs."""
    g = torch.Generator().manual_seed(rows * 7 + cols)
    calls = []
    for c in range(3):
        x = torch.randn(rows, cols, generator=g) * (1 + c)
        x[:, cols // 3] *= 50
        calls.append(x.to(DT[dt]))
    acc_max = torch.zeros(cols, dtype=torch.float64, device=DEV)
    acc_sum = torch.zeros(cols, dtype=torch.float64, device=DEV)
    run = torch.empty(cols, dtype=DT[dt], device=DEV)
    for c, x in enumerate(calls):
        out = qdm.ops.colstats(x.to(DEV), out_max=run, running=c > 0, acc_maxsum=acc_max, acc_abssum=acc_sum)
        assert out is run
        # the per-call statistic alone, through the same kernel
        assert_bit_equal(qdm.ops.colstats(x.to(DEV)), O.hook_colabsmax(x), f"per-call max {c}")
    per_call = [O.hook_colabsmax(x) for x in calls]
    assert_bit_equal(run, torch.stack(per_call).amax(0), "running max")
    want_max = torch.stack(per_call).double().sum(0)
    assert torch.equal(acc_max.cpu(), want_max), "fp64 sum of per-call maxima must be exact"
    want_sum = sum(x.abs().double().sum(0) for x in calls)
    assert ((acc_sum.cpu() - want_sum).abs() / want_sum.clamp_min(1e-30)).max().item() < 1e-5
    # subsets of the outputs: only the |x| sum; only the max-sum
    only = torch.zeros(cols, dtype=torch.float64, device=DEV)
    assert qdm.ops.colstats(calls[0].to(DEV), acc_abssum=only) is None
    assert ((only.cpu() - calls[0].abs().double().sum(0)).abs() / want_sum.clamp_min(1e-30)).max().item() < 1e-5
    assert torch.equal(only.float().cpu(), qdm.ops.colabssum(calls[0].to(DEV)).cpu())   # same tree as qdm_colabssum
    only.zero_()
    qdm.ops.colstats(calls[1].to(DEV), acc_maxsum=only)
    assert torch.equal(only.cpu(), per_call[1].double())


def test_colstats_strided_and_bad_inputs(qdm):
    x = torch.randn(64, 4, 96, generator=torch.Generator().manual_seed(5)).half()
    assert_bit_equal(qdm.ops.colstats(x.to(DEV)), O.hook_colabsmax(x), "3-D input")
    xs = x.to(DEV)[:, :, :40]                      # row stride 96, 40 columns: ld != cols
    assert_bit_equal(qdm.ops.colstats(xs), O.hook_colabsmax(x[:, :, :40]), "strided rows")
    with pytest.raises(ValueError):
        qdm.ops.colstats(torch.zeros(0, 8, dtype=torch.float16, device=DEV))
    with pytest.raises(ValueError):
        qdm.ops.colstats(x.to(DEV), acc_maxsum=torch.zeros(96, dtype=torch.float32, device=DEV))
    with pytest.raises(ValueError):
        qdm.ops.colstats(x.to(DEV), running=True)
    with pytest.raises(RuntimeError):
        qdm.ops.colstats(x)
