"""GPU parity of the same-input fusion (fused_utils.py; the reference's utils/fused_utils.py:45-163) and of qdm_geglu.

  * every launch shape the fusion creates for the three denoisers (attn1 q/k/v, attn2 k/v of all blocks, the grouped
    time_emb_proj / AdaLN launches -- N up to 1 104 128) against F.linear on the fake-quant weight in fp32, <= 1e-2;
  * the fused module's output against its members' outputs, and its packed tensors against the members' (bit-equal);
  * the fused denoiser against the unfused one over a multi-step CFG loop, <= 1e-2 (north star: denoised latents);
  * qdm_geglu against torch's `h * F.gelu(gate)` (the two ops it replaces), f16 / bf16.
"""
import importlib

import pytest
import torch
import torch.nn.functional as F

from _util import DT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2
PKG = "quantization---diffusion-models_b200"

shapes = importlib.import_module(PKG + ".shapes")


def _new_shapes():
    base = set()
    for f in (shapes.sd15_unet_linears, shapes.sdxl_unet_linears, shapes.sd35_mmdit_linears):
        base |= {(m, n, k) for _, m, n, k, _ in f()}
    out = []
    for model, f in (("sd15", shapes.sd15_unet_linears_fused), ("sdxl", shapes.sdxl_unet_linears_fused), ("sd35", shapes.sd35_mmdit_linears_fused)):
        for e in f():
            s = (e[1], e[2], e[3])
            if s not in base and all(s != o[1:4] for o in out):
                out.append((model,) + s + (e[5],))
    return out


FUSED_SHAPES = _new_shapes()


def test_fused_inventory_is_what_the_docs_say():
    got = {c[:4] for c in FUSED_SHAPES}
    assert {("sd15", 65536, 960, 320), ("sd15", 1232, 24960, 768), ("sd15", 16, 17600, 1280), ("sd35", 1, 1104128, 2432)} <= got


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("model,M,N,K,parts", FUSED_SHAPES, ids=[f"{c[0]}-{c[1]}x{c[2]}x{c[3]}" for c in FUSED_SHAPES])
def test_w4a16_fused_launch_shape(qdm, model, M, N, K, parts, dt):
    """The launch as the model builds it: every member Linear quantised on its own, packed tensors concatenated along N."""
    assert not torch.backends.cuda.matmul.allow_tf32
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + 7 * K)
    grp = shapes.group_for(K)
    x = torch.randn(M, K, generator=g, device=DEV, dtype=DT[dt])
    linear = importlib.import_module(PKG + ".linear")
    fu = importlib.import_module(PKG + ".fused_utils")
    mems = []
    for n in parts:
        lin = torch.nn.Linear(K, n, bias=True, device=DEV, dtype=DT[dt])
        lin.weight.data = (torch.randn(n, K, generator=g, device=DEV) * 0.05).to(DT[dt])
        lin.bias.data = torch.randn(n, generator=g, device=DEV).to(DT[dt])
        mems.append(linear.WQLinear_GEMM.from_linear(lin, 4, grp))
        del lin
    mod = fu.fuse_linears(mems)
    assert mod.out_features == N
    b = mod.bias
    del mems
    y = mod(x)
    variant, tile = qdm.ops.gemm_last_variant()
    assert variant in ({"skinny", "smallm"} if M <= 32 else {"pair", "bstat", "streamk", "ts"}), (variant, tile)
    assert y.shape == (M, N) and y.dtype == DT[dt] and torch.isfinite(y).all()
    # reference in column slabs: dequantise a slab of the packed tensors, F.linear in fp32
    worst, ref_max = 0.0, 0.0
    slab = 8192 * 8
    for lo in range(0, N, slab):
        hi = min(N, lo + slab)
        dq = qdm.ops.dequant_awq(mod.qweight[:, lo // 8:hi // 8].contiguous(), mod.qzeros[:, lo // 8:hi // 8].contiguous(),
                                 mod.scales[:, lo:hi].contiguous(), grp).t()   # [n_slab, K]
        for r in range(0, M, 8192):
            ref = x[r:r + 8192].float() @ dq.float().t() + b[lo:hi].float()
            worst = max(worst, (y[r:r + 8192, lo:hi].float() - ref).abs().max().item())
            ref_max = max(ref_max, ref.abs().max().item())
    assert worst / ref_max <= TOL, (variant, tile, worst / ref_max)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,K,parts,bias", [(4096, 1280, (1280, 1280, 1280), False), (1232, 768, (320, 320, 640, 640, 1280, 1280), False),
                                            (16, 1280, (320, 640, 1280, 320), True), (300, 320, (320, 320, 320), True)])
def test_fuse_linears_equals_members(qdm, M, K, parts, bias, dt):
    linear = importlib.import_module(PKG + ".linear")
    fu = importlib.import_module(PKG + ".fused_utils")
    g = torch.Generator(device=DEV).manual_seed(5 + M + K)
    grp = shapes.group_for(K)
    mems = []
    for i, n in enumerate(parts):
        lin = torch.nn.Linear(K, n, bias=bias and i != 1, device=DEV, dtype=DT[dt])   # one member without a bias
        lin.weight.data = (torch.randn(n, K, generator=g, device=DEV) * 0.05).to(DT[dt])
        mems.append(linear.WQLinear_GEMM.from_linear(lin, 4, grp))
    fused = fu.fuse_qkv(None, *mems) if len(mems) == 3 else fu.fuse_linears(mems)
    assert fused.out_features == sum(parts) and fused.split_sizes == list(parts)
    assert torch.equal(fused.dequantize(), torch.cat([m.dequantize() for m in mems]))        # packed tensors concatenate code for code
    x = torch.randn(2, M // 2, K, generator=g, device=DEV, dtype=DT[dt])                     # 3-D input like the model's
    y = fused(x)
    want = torch.cat([m(x) for m in mems], dim=-1)
    assert y.shape == want.shape
    err = ((y.float() - want.float()).abs().max() / want.float().abs().max()).item()
    # same codes, fp32 accumulation; the kernels may differ in summation order and in where the scale is applied (per weight,
    # rounded to the 16-bit type, or per group in fp32 -- the M <= 8 kernel): one ulp of the type at the largest magnitude
    assert err <= (2e-3 if dt == "f16" else 8e-3), err
    with pytest.raises(ValueError):
        other = linear.WQLinear_GEMM(4, grp, K * 2, 64, False, DEV, DT[dt])
        fu.fuse_linears([mems[0], other])


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("M,K,parts", [(4096, 1280, (1280, 1280, 1280)), (1232, 768, (320, 640, 1280, 320)), (300, 320, (320, 320, 320))])
def test_fuse_w8a8_equals_members(qdm, M, K, parts, dt):
    """W8A8Linear members: rows of int8 codes and row scales concatenate; the shared input is quantised once.  Integer
    accumulation is exact, so the fused output equals the members' bit for bit."""
    linear = importlib.import_module(PKG + ".linear")
    fu = importlib.import_module(PKG + ".fused_utils")
    g = torch.Generator(device=DEV).manual_seed(9 + M + K)
    smooth = (torch.rand(K, generator=g, device=DEV) + 0.5).to(DT[dt])
    mems = []
    for i, n in enumerate(parts):
        lin = torch.nn.Linear(K, n, bias=i != 1, device=DEV, dtype=DT[dt])
        lin.weight.data = (torch.randn(n, K, generator=g, device=DEV) * 0.05).to(DT[dt])
        mems.append(linear.W8A8Linear.from_float(lin, smooth=smooth))
    fused = fu.fuse_linears(mems)
    x = torch.randn(2, M // 2, K, generator=g, device=DEV, dtype=DT[dt])
    qdm.ops.launch_count(reset=True)
    y = fused(x)
    assert qdm.ops.launch_count() == 2                         # one per-token quantiser + one GEMM
    want = torch.cat([m(x) for m in mems], dim=-1)
    assert torch.equal(y, want)
    mems[1].smooth = smooth * 2
    with pytest.raises(ValueError):
        fu.fuse_linears(mems)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("shape", [(4096, 2560), (3, 77, 5120), (1, 16), (1000, 10240), (65536, 2560)])
def test_geglu_vs_torch(qdm, shape, dt):
    g = torch.Generator(device=DEV).manual_seed(sum(shape))
    x = (torch.randn(*shape, generator=g, device=DEV) * 3).to(DT[dt])
    x.view(-1)[:4] = torch.tensor([0.0, -0.0, 60000.0 if dt == "f16" else 1e30, -8.0], device=DEV).to(DT[dt])
    y = qdm.ops.geglu(x)
    h, gate = x.chunk(2, dim=-1)
    ref = h * F.gelu(gate)                                     # the two torch ops of skeletons.GEGLU / diffusers GEGLU
    assert y.shape == ref.shape and y.dtype == ref.dtype
    assert torch.equal(torch.isfinite(y), torch.isfinite(ref))
    fin = torch.isfinite(ref)
    # same fp32 formula, same two roundings: where erff's last bit moves the rounded gelu by one 16-bit ulp, the product moves
    # by a relative 2^-10 (f16) / 2^-7 (bf16) plus its own rounding; everywhere else the results are bit-equal
    diff = (y.float() - ref.float()).abs()[fin]
    bound = ref.float().abs()[fin] * (2.0 ** -9 if dt == "f16" else 2.0 ** -6) + 1e-7
    assert (diff <= bound).all(), (diff / bound).max().item()
    assert (y[fin] == ref[fin]).float().mean().item() >= 0.98
    with pytest.raises(ValueError):
        qdm.ops.geglu(torch.zeros(4, 24, device=DEV, dtype=DT[dt]))


def test_fused_w8a8_denoiser_matches_unfused(qdm):
    """BASELINE config 3 in small: SmoothQuant alpha = 0.5 + real W8A8 modules on the SDXL skeleton, fused against unfused."""
    M = importlib.import_module(PKG + ".models")
    model = M.StableDiffusionXL.from_skeleton(device=DEV, channels=(320, 640), depth=(0, 2), latent_size=32)
    lat = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(11)).half().to(DEV)
    model.calib_samples = model.default_calib_samples(1, 2)
    model.quantize(quant_config={"w_bit": 8, "version": "w8a8"}, quantType="sq", alpha=0.5, calib_num_infer_steps=1)
    assert any(type(m).__name__ == "W8A8Linear" for m in model.denoiser().modules())
    ref = model.generate(["a", "b"], lat=lat, num_inference_steps=3)
    done = model.fuse_layers()
    assert done["self_qkv"] > 0 and done["context_kv"] > 0 and done["time_emb_proj"] > 0
    out = model.generate(["a", "b"], lat=lat, num_inference_steps=3, fuse_layers=True)
    err = ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
    assert err <= TOL, err            # the GEMMs are bit-equal; qdm_geglu may differ from gelu + mul in the last bit


@pytest.mark.parametrize("kind", ["sd15", "sdxl", "sd35"])
def test_fused_denoiser_matches_unfused(qdm, kind):
    """fuse_layers(): the same packed model with the same-input projections as one launch each and qdm_geglu, against the
    unfused model, over a CFG denoise loop (eager and graph-replayed)."""
    M = importlib.import_module(PKG + ".models")
    if kind == "sd15":
        model = M.StableDiffusion1_x.from_skeleton(device=DEV, latent_size=32)
        lat = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(11)).half().to(DEV)
    elif kind == "sdxl":
        model = M.StableDiffusionXL.from_skeleton(device=DEV, latent_size=32)
        lat = torch.randn(2, 4, 32, 32, generator=torch.Generator().manual_seed(11)).half().to(DEV)
    else:
        model = M.StableDiffusion3_5.from_skeleton(device=DEV, layers=3, latent_size=32)
        lat = torch.randn(2, 16, 32, 32, generator=torch.Generator().manual_seed(11)).half().to(DEV)
    model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq")
    qdm.ops.launch_count(reset=True)
    ref = model.generate(["a", "b"], lat=lat, num_inference_steps=3)
    n_unfused = qdm.ops.launch_count()
    done = model.fuse_layers()
    assert done["self_qkv"] > 0
    if kind == "sd35":
        assert done["adaln"] >= 7 and done["context_kv"] == 0
    else:
        assert done["context_kv"] > 0 and done["time_emb_proj"] > 0 and done["geglu"] > 0
    names = set(model.denoiser().state_dict().keys())
    qdm.ops.launch_count(reset=True)
    out = model.generate(["a", "b"], lat=lat, num_inference_steps=3, fuse_layers=True)
    n_fused = qdm.ops.launch_count()
    assert n_fused < n_unfused, (n_fused, n_unfused)
    assert set(model.denoiser().state_dict().keys()) == names          # fused copies are not part of the state dict
    err = ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item()
    assert err <= TOL, err
    out_g = model.generate(["a", "b"], lat=lat, num_inference_steps=3, fuse_layers=True, cuda_graph=True)
    err_g = ((out_g.float() - out.float()).abs().max() / out.float().abs().max()).item()
    assert err_g <= TOL, err_g
    model.unfuse_layers()
    again = model.generate(["a", "b"], lat=lat, num_inference_steps=3)
    assert torch.equal(again, ref)
