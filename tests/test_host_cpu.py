"""CPU tests of the HOST logic around the kernels: module swap (incl. pointwise convolutions), packed checkpoint,
fused hook statistic plumbing, and the block-sharded AWQ search at world size 2 (gloo) against world size 1.

The product has no CPU compute path, so these tests run the host code with `ops` patched by the oracle
(tests/_oracle_ops.py, test infrastructure).  The kernels themselves are checked on the GPU (`-m gpu`)."""
import importlib
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _oracle_ops import PKG, patched_ops

ARCH = dict(channels=(64, 128), depth=(1, 1), ctx_dim=64, heads=2, latent_size=16)


def tiny_sd15():
    M = importlib.import_module(PKG + ".models")
    return M, M.StableDiffusion1_x.from_skeleton(device="cpu", **ARCH)


def packed_state(model):
    """every integer / scale buffer of the swapped modules, keyed by name (what must be identical across world sizes)"""
    return {k: v.clone() for k, v in model.denoiser().state_dict().items()
            if k.endswith(("qweight", "qzeros", "scales", "w_scales"))}


def test_ops_reject_cpu_tensors():
    """no CPU fallback in the product: the un-patched front end refuses CPU tensors before touching the library"""
    ops = importlib.import_module(PKG + ".ops")
    x = torch.zeros(4, 8, dtype=torch.float16)
    for call in (lambda: ops.colabsmax(x), lambda: ops.colstats(x), lambda: ops.quant_group(x, 8),
                 lambda: ops.gemm_f16(x, x), lambda: ops.quant_pack_awq(x, 8)):
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            call()


@pytest.mark.parametrize("version,inner", [("gemm", "WQLinear_GEMM"), ("w8a8", "W8A8Linear")])
def test_swap_pointwise_convs_and_checkpoint(version, inner, tmp_path):
    with patched_ops():
        M, model = tiny_sd15()
        lat = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(1)).half()
        fp = model.generate(["a", "b"], lat=lat, num_inference_steps=1).float()
        n_pw = sum(1 for m in model.denoiser().modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (1, 1))
        n_33 = sum(1 for m in model.denoiser().modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3))
        model.quantize(quant_config={"q_group_size": 64, "w_bit": 4 if version == "gemm" else 8, "version": version},
                       quantType="awq")
        mods = [m for m in model.denoiser().modules() if type(m).__name__ == "QConv1x1"]
        assert len(mods) == n_pw > 0 and all(type(m.inner).__name__ == inner for m in mods)
        # 3x3 / pad 1 convolutions (stride 1 and the stride-2 down-samplers) with C % 64 == 0 carry packed int4 weights
        # too (implicit-GEMM kernel c); conv_in (C = 4) and conv_out (N = 4) keep fake-quant weights
        n_q3 = sum(1 for m in model.denoiser().modules() if type(m).__name__ == "QConv3x3")
        n_fake = sum(1 for m in model.denoiser().modules() if type(m).__name__ == "WxAxConv2d")
        assert n_q3 + n_fake == n_33 and n_fake >= 2
        assert (n_q3 > 0) == (version == "gemm")
        assert not any(isinstance(m, (torch.nn.Linear, torch.nn.Conv2d)) for m in model.denoiser().modules())
        out = model.generate(["a", "b"], lat=lat, num_inference_steps=1)
        assert torch.isfinite(out).all()
        assert ((out.float() - fp).abs().max() / fp.abs().max()).item() < (0.5 if version == "gemm" else 0.1)
        model.save_quantized(str(tmp_path))
        # the checkpoint is a safetensors file whose metadata carries the Hugging Face quantization_config (models/base.py:530-582)
        import json
        from safetensors import safe_open
        with safe_open(str(tmp_path / "model.safetensors"), framework="pt") as f:
            qc = json.loads(f.metadata()["quantization_config"])
            keys = set(f.keys())
        assert qc["quant_method"] == "awq" and qc["bits"] == (4 if version == "gemm" else 8) and "group_size" in qc
        assert keys == set(model.denoiser().state_dict().keys())
        again = M.StableDiffusion1_x.from_quantized(str(tmp_path), device="cpu")
        for k, v in packed_state(model).items():
            assert torch.equal(v, again.denoiser().state_dict()[k]), k
        assert torch.equal(again.generate(["a", "b"], lat=lat, num_inference_steps=1), out)


def test_sq_fused_hook_plumbing():
    """quantize('sq') with the fused one-pass hook == the per-call hook (oracle arithmetic: identical or 1 ulp)"""
    logs = []
    with patched_ops():
        for fused in (False, True):
            M, model = tiny_sd15()
            model.calib_samples = model.default_calib_samples(1, 2)
            model.quantize(quant_config={"w_bit": 8, "version": "w8a8"}, quantType="sq", alpha=0.5, calib_num_infer_steps=2,
                           fused_stats=fused)
            logs.append(model.quantizer.smooth_log)
    assert logs[0].keys() == logs[1].keys() and len(logs[0]) > 0
    for k in logs[0]:
        for a, b in zip(logs[0][k], logs[1][k]):
            assert ((a.float() - b.float()).abs() <= b.float().abs() * 2 ** -9).all()


def _calibrated(shard, kind="sd15"):
    M, model = tiny_model(kind)
    model.calib_steps = 2
    model.quantize(quant_config={"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "gemm"}, quantType="awq",
                   calibrate=True, shard=shard)
    return model


def _shard_worker(rank, world, port, q, kind):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with patched_ops():
        model = _calibrated((rank, world), kind)
    q.put((rank, {k: v.numpy() for k, v in packed_state(model).items()}, len(model.quantizer.search_log)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["sd15", "sd35"])
def test_sharded_awq_search_world2_equals_world1(kind):
    """blocks are searched on the rank that owns them, ONE gather exchanges {scales, clip}, every rank
    applies the identical list -> the packed codes / zeros / scales are bit-identical for world 1 and world 2, on
    both ranks, and each rank searched only its own blocks."""
    with patched_ops():
        single = _calibrated(None, kind)
    want = packed_state(single)
    n_groups = len(single.quantizer.search_log)
    assert n_groups >= 3 * len(single.get_search_blocks())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000 + (7 if kind == "sd35" else 0)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][2] + res[1][2] == n_groups and 0 < res[0][2] < n_groups
    for rank, state, _ in res:
        assert state.keys() == want.keys()
        for k, v in want.items():
            assert (torch.from_numpy(state[k]) == v).all(), f"rank {rank}: {k} differs from the single-rank result"


def _hook_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cd = importlib.import_module(PKG + ".calib_data")
    d = importlib.import_module(PKG + ".dist")
    g = torch.Generator().manual_seed(7)
    calls = [torch.randn(12, 3, 40, generator=g).half() * (1 + c) for c in range(6)]
    with patched_ops():
        hooks = {"blk": {"to_q": cd.Fused_Mean_Max_Activation_Hook(want_abssum=True), "ff": cd.Fused_Mean_Max_Activation_Hook()}}
        for c in range(rank, 6, world):          # this rank's share of the calibration calls
            hooks["blk"]["to_q"](None, (calls[c],), None)
            hooks["blk"]["ff"](None, (calls[c][..., :24].contiguous(),), None)
        d.allreduce_hook_stats(hooks)
    h = hooks["blk"]["to_q"]
    q.put((rank, h.step, h.rows, h.max_scales.mean_over_calls().numpy(), h.x_mean().numpy(), h.running_max.numpy(),
           hooks["blk"]["ff"].max_scales.mean_over_calls().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_hook_stats_allreduce_world2_is_exact():
    """SURVEY 8(e), SmoothQuant statistics under data-parallel calibration: after one all_reduce of the fp64 accumulators
    both ranks hold exactly the single-process statistic (the sum of fp16 maxima is exact in fp64, so order-free)."""
    cd = importlib.import_module(PKG + ".calib_data")
    g = torch.Generator().manual_seed(7)
    calls = [torch.randn(12, 3, 40, generator=g).half() * (1 + c) for c in range(6)]
    with patched_ops():
        ref, ref_ff = cd.Fused_Mean_Max_Activation_Hook(want_abssum=True), cd.Fused_Mean_Max_Activation_Hook()
        for x in calls:
            ref(None, (x,), None)
            ref_ff(None, (x[..., :24].contiguous(),), None)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_hook_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, step, rows, mean_max, x_mean, rmax, ff_mean in res:
        assert step == 6 and rows == 6 * 36
        assert (torch.from_numpy(mean_max) == ref.max_scales.mean_over_calls()).all()
        assert (torch.from_numpy(rmax) == ref.running_max).all()
        assert (torch.from_numpy(ff_mean) == ref_ff.max_scales.mean_over_calls()).all()
        # |x| sums: per-call fp32 sums folded in fp64 -- the same addends on any rank split, added in a different
        # order in fp64: equal to the last bit after rounding to fp16
        assert (torch.from_numpy(x_mean) == ref.x_mean()).all()


def test_conv_helpers():
    L = importlib.import_module(PKG + ".linear")
    C = torch.nn.Conv2d
    assert L.is_pointwise_conv(C(8, 16, 1)) and not L.is_pointwise_conv(C(8, 16, 1, stride=2)) and not L.is_pointwise_conv(C(8, 16, 3, padding=1))
    assert L.is_conv3x3_gemm(C(64, 16, 3, padding=1)) and L.is_conv3x3_gemm(C(64, 16, 3, padding=1, stride=2))   # down-samplers too
    assert not L.is_conv3x3_gemm(C(64, 16, 3, padding=1, stride=3)) and not L.is_conv3x3_gemm(C(64, 16, 3, padding=1, stride=(2, 1)))
    assert not L.is_conv3x3_gemm(C(4, 320, 3, padding=1)) and not L.is_conv3x3_gemm(C(64, 16, 3)) and not L.is_conv3x3_gemm(C(64, 4, 3, padding=1))
    assert not L.is_conv3x3_gemm(C(64, 64, 3, padding=1, groups=2)) and not L.is_conv3x3_gemm(C(64, 64, 3, padding="same"))
    assert [L.conv_group(9 * c, 128) for c in (64, 320, 640, 960, 1280, 1920, 2560)] == [64, 64, 128, 64, 128, 128, 128]
    x = torch.randn(2, 8, 3, 5)
    t = L.nchw_as_tokens(x)
    assert t.shape == (30, 8) and torch.equal(L.tokens_as_nchw(t, 2, 3, 5), x)
    xc = x.contiguous(memory_format=torch.channels_last)
    assert L.nchw_as_tokens(xc).data_ptr() == xc.data_ptr()          # a view, no copy
    assert L.tokens_as_nchw(t, 2, 3, 5).is_contiguous(memory_format=torch.channels_last)


# ------------------------------------------------------------------ SDXL / SD3.5 adapters (host logic, oracle-backed ops)
def tiny_model(kind, dtype=torch.float16):
    M = importlib.import_module(PKG + ".models")
    if kind == "sd35":
        return M, M.StableDiffusion3_5.from_skeleton(device="cpu", dtype=dtype, layers=2, dim=128, heads=2, in_ch=4, patch=2,
                                                     ctx_in=64, pooled=32, latent_size=8)
    if kind == "sdxl":
        return M, M.StableDiffusionXL.from_skeleton(device="cpu", dtype=dtype, channels=(64, 128), depth=(0, 1), ctx_dim=64,
                                                    head_dim=32, add_embed_in=96, latent_size=16)
    return M, M.StableDiffusion1_x.from_skeleton(device="cpu", dtype=dtype, **ARCH)


@pytest.mark.parametrize("kind", ["sd15", "sdxl", "sd35"])
def test_awq_scale_fold_preserves_the_fp_function(kind):
    """apply_scale (quantize/scale.py:37-84; AdaLN variant for the MMDiT) only moves a per-channel factor from the
    weights of a scaling group into the op before it: with NO quantisation the denoiser must compute the same function.
    Run in fp32 so that the check is tight; the scales come from the real search (oracle-backed ops)."""
    with patched_ops():
        M, model = tiny_model(kind, torch.float32)
        model.calib_steps = 1
        lat = torch.randn(2, model.pipeline.latent_channels, model.pipeline.latent_size, model.pipeline.latent_size,
                          generator=torch.Generator().manual_seed(5))
        before = model.generate(["a", "b"], lat=lat, num_inference_steps=1)
        Q = importlib.import_module(PKG + ".quantizer").AwqQuantizer
        quant = Q(model, None, None, group_size=64, zero_point=True, version="gemm", calibrate=True, apply_clip=False,
                  quantUnet=model.pipeline.unet is not None, quantTransformer=model.pipeline.transformer is not None)
        results = quant.search()
        assert len(results) == len(model.get_search_blocks()) > 0
        n_groups = sum(len(r["scales"]) for r in results.values())
        assert n_groups >= 3 * len(results)
        assert any((s != 1).any() for r in results.values() for _, _, s in r["scales"])      # the fold is not a no-op
        quant.apply_search_results(results)
        after = model.generate(["a", "b"], lat=lat, num_inference_steps=1)
    assert torch.isfinite(after).all()
    assert ((after - before).abs().max() / before.abs().max()).item() < 1e-4


@pytest.mark.parametrize("kind", ["sdxl", "sd35"])
def test_quantize_gemm_and_w8a8_on_other_adapters(kind, tmp_path):
    """quantize('awq', calibrate=True, version='gemm') and quantize('sq', version='w8a8') on the SDXL / SD3.5 skeletons:
    every Linear swapped, finite latents close to the FP model, packed checkpoint round trip."""
    with patched_ops():
        M, model = tiny_model(kind)
        model.calib_steps = 1
        lat = torch.randn(2, model.pipeline.latent_channels, model.pipeline.latent_size, model.pipeline.latent_size,
                          generator=torch.Generator().manual_seed(6)).half()
        fp = model.generate(["a", "b"], lat=lat, num_inference_steps=1).float()
        model.quantize(quant_config={"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "gemm"}, quantType="awq",
                       calibrate=True)
        kinds = {type(m).__name__ for m in model.denoiser().modules()}
        assert "WQLinear_GEMM" in kinds
        out = model.generate(["a", "b"], lat=lat, num_inference_steps=1)
        assert torch.isfinite(out).all()
        assert ((out.float() - fp).abs().max() / fp.abs().max()).item() < 0.5
        model.save_quantized(str(tmp_path))
        cls = type(model)
        again = cls.from_quantized(str(tmp_path), device="cpu")
        assert torch.equal(again.generate(["a", "b"], lat=lat, num_inference_steps=1), out)
        M2, m8 = tiny_model(kind)
        m8.calib_samples = m8.default_calib_samples(1, 2)
        if kind == "sd35":   # SmoothQuant groups exist for UNet BasicTransformerBlocks only, as in the reference
            with pytest.raises(NotImplementedError):
                m8.quantize(quant_config={"w_bit": 8, "version": "w8a8"}, quantType="sq", alpha=0.5, calib_num_infer_steps=1)
            return
        m8.quantize(quant_config={"w_bit": 8, "version": "w8a8"}, quantType="sq", alpha=0.5, calib_num_infer_steps=1,
                    fused_stats=True)
        assert any(type(m).__name__ == "W8A8Linear" for m in m8.denoiser().modules())
        q8 = m8.generate(["a", "b"], lat=lat, num_inference_steps=1).float()
        assert torch.isfinite(q8).all() and ((q8 - fp).abs().max() / fp.abs().max()).item() < 0.2


def test_packing_mirrors_on_reference_fixture():
    """packing_utils / quant_utils mirrors (names and argument meaning of utils/packing_utils.py, utils/quant_utils.py)
    against the reference-generated layout fixture: pack, unpack in packed order, order reversal, dequantize_gemm."""
    from _util import Golden
    pu = importlib.import_module(PKG + ".packing_utils")
    qu = importlib.import_module(PKG + ".quant_utils")
    g = Golden("awq_layout.npz")
    codes, zeros = g.get("codes"), g.get("zeros")
    qweight, qzeros, scales, deq = g.get("qweight"), g.get("qzeros"), g.get("scales"), g.get("deq")
    gs = int(g.get("group"))
    with patched_ops():
        assert torch.equal(qu.pack_awq(codes), qweight) and torch.equal(qu.pack_awq(zeros), qzeros)
        assert torch.equal(qu.unpack_awq(qweight).to(torch.int32), codes)
        iw, iz = pu.unpack_awq(qweight, qzeros, 4)                 # still in packed nibble order, like the reference
        assert not torch.equal(iw.to(torch.int32), codes)
        iw, iz = pu.reverse_awq_order(iw, iz, 4)
        assert torch.equal(iw.to(torch.int32) & 0xF, codes) and torch.equal(iz.to(torch.int32) & 0xF, zeros)
        assert torch.equal(pu.dequantize_gemm(qweight, qzeros, scales, 4, gs), deq)
        assert torch.equal(qu.dequantize(codes, scales, zeros, gs), deq.to(torch.float16))
        with pytest.raises(NotImplementedError):
            pu.dequantize_gemm(qweight, qzeros, scales, 8, gs)
    x = torch.arange(32).view(2, 16)
    assert torch.equal(qu.apply_order(qu.apply_order(x, "column", qu.AWQ_PACK_ORDER), "column", qu.REVERSE_AWQ_PACK_ORDER), x)
    with pytest.raises(ValueError):
        qu.apply_order(x, "diagonal")


def test_awq_search_host_logic_on_reference_fixture():
    """AwqQuantizer._search_best_scale / _compute_best_scale / _compute_best_clip (the host orchestration: statistics,
    20-point ratio grid, restore of the weights, first-minimum argmin, clip levels) against the reference-generated
    search fixture, with the kernels stood in by the oracle."""
    from _util import Golden
    import oracle.qdm_oracle as O
    Q = importlib.import_module(PKG + ".quantizer").AwqQuantizer

    class Cat(torch.nn.Module):
        def __init__(self, ls):
            super().__init__()
            self.ls = torch.nn.ModuleList(ls)

        def forward(self, x):
            return torch.cat([l(x) for l in self.ls], dim=-1)

    g = Golden("awq_search.npz")
    with patched_ops():
        for tag, dt, group, zp in g.cases():
            x = g.get(tag + "_x")
            lins = []
            for i in range(3):
                w, b = g.get(f"{tag}_w{i}"), g.get(f"{tag}_b{i}")
                l = torch.nn.Linear(w.shape[1], w.shape[0], bias=True)
                l.weight.data, l.bias.data = w.clone(), b.clone()
                lins.append(l)
            block = Cat(lins)
            q = Q(None, group_size=int(group), zero_point=bool(int(zp)))
            before = [l.weight.data.clone() for l in lins]
            prev, names, best = q._search_best_scale(block, lins[0], lins, x, module2inspect=block, kwargs={})
            assert names == ("ls.0", "ls.1", "ls.2") and prev == "ls.0"
            for l, wb in zip(lins, before):
                assert torch.equal(l.weight.data, wb)                # originals restored
            want = g.get(tag + "_best")
            if not torch.equal(best, want):                          # near-tie between two ratios: loss within 0.5 %
                ws, bs = [g.get(f"{tag}_w{i}") for i in range(3)], [g.get(f"{tag}_b{i}") for i in range(3)]
                fwd = lambda wl: torch.cat([torch.nn.functional.linear(x, w, b) for w, b in zip(wl, bs)], dim=-1)
                _, _, hist = O.awq_search_scale(x, ws, fwd, int(group), bool(int(zp)))
                assert hist[int(round(q.last_best_ratio * 20))] <= min(hist) * 1.005
            clip, ref_clip = q._compute_best_clip(lins[0].weight.data, x), g.get(tag + "_clip")
            assert clip.shape == ref_clip.shape
            assert (clip == ref_clip).float().mean().item() >= 0.97


def test_calibrate_with_modules_to_not_convert_scales_once():
    """a scaling-group member excluded from quantisation (modules_to_not_convert=['to_k']) is still scaled by
    apply_scale during the clip phase; it must be rolled back like every other Linear so that apply_search_results
    scales it exactly ONCE.  With scales applied once and no quantisation of that layer, weight == original * s."""
    with patched_ops():
        M, model = tiny_sd15()
        model.calib_steps = 1
        blocks = model.get_search_blocks()
        bname = next(iter(blocks))
        to_k = blocks[bname].attn1.to_k
        w0 = to_k.weight.data.clone()
        Q = importlib.import_module(PKG + ".quantizer").AwqQuantizer
        quant = Q(model, None, None, group_size=64, zero_point=True, version="gemm", calibrate=True, apply_clip=True,
                  modules_to_not_convert=["to_k"])
        results = quant.search()
        assert torch.equal(to_k.weight.data, w0)                      # search leaves every weight as it found it
        s = next(sc for prev, names, sc in results[bname]["scales"] if "attn1.to_k" in names)
        quant.apply_search_results(results)
        assert torch.equal(to_k.weight.data, w0 * s.view(1, -1).to(w0.dtype))
        assert all("to_k" not in name for name, _ in results[bname]["clip"])


def test_non_tileable_groups_fall_back_to_fake_quant():
    """group_size = -1 (per-channel: g = K = 64 * 5 for K = 320) or 192 cannot run on the W4A16 kernel
    (g must be 64 * 2^j): the swap must keep those layers on the fake-quant path instead of building a module whose
    forward raises."""
    L = importlib.import_module(PKG + ".linear")
    assert L.w4a16_kernel_ok(320, 320, 64) and L.w4a16_kernel_ok(2432, 9728, 128) and L.w4a16_kernel_ok(1024, 8, 256)
    assert not L.w4a16_kernel_ok(320, 320, 320) and not L.w4a16_kernel_ok(384, 64, 192) and not L.w4a16_kernel_ok(64, 12, 64)
    assert not L.w4a16_kernel_ok(96, 64, 96) and not L.w4a16_kernel_ok(128, 64, 0)
    assert L.w8a8_kernel_ok(320, 8) and not L.w8a8_kernel_ok(72, 8) and not L.w8a8_kernel_ok(64, 12)
    with patched_ops():
        M = importlib.import_module(PKG + ".models")
        model = M.StableDiffusion1_x.from_skeleton(device="cpu", channels=(320,), depth=(1,), ctx_dim=64, heads=2, latent_size=8)
        model.quantize(quant_config={"q_group_size": -1, "w_bit": 4, "version": "gemm"}, quantType="awq")
        kinds = [type(m).__name__ for m in model.denoiser().modules()]
        assert "WxAxLinear" in kinds                                   # K = 320 per-channel: fake-quant fallback
        lat = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(2)).half()
        assert torch.isfinite(model.generate(["p"], lat=lat, num_inference_steps=1)).all()


def test_fake_act_checkpoint_keeps_activation_quantisers(tmp_path):
    """quantize_act / a_bit / per-group conv activations are constructor arguments, not state: the packed
    checkpoint must carry them, or a model quantised with activation quantisation ON reloads with it OFF."""
    with patched_ops():
        M, model = tiny_sd15()
        lat = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(3)).half()
        model.quantize(quant_config={"q_group_size": 64, "w_bit": 8, "a_bit": 8, "version": "fake_act", "quantize_act": True,
                                     "act_quant_conv_type": "per_group", "act_quant_conv_group_size": 4}, quantType="awq")
        convs = [m for m in model.denoiser().modules() if type(m).__name__ == "WxAxConv2d"]
        assert convs and all(c.quantise_act and c.act_quant_name == "per_group" and c.a_gs == 4 and c.n_bits_A == 8 for c in convs)
        out = model.generate(["p"], lat=lat, num_inference_steps=1)
        model.save_quantized(str(tmp_path))
        again = M.StableDiffusion1_x.from_quantized(str(tmp_path), device="cpu")
        convs2 = [m for m in again.denoiser().modules() if type(m).__name__ == "WxAxConv2d"]
        assert len(convs2) == len(convs)
        assert all(c.quantise_act and c.act_quant_name == "per_group" and c.a_gs == 4 and c.n_bits_A == 8 for c in convs2)
        lins2 = [m for m in again.denoiser().modules() if type(m).__name__ == "WxAxLinear"]
        assert lins2 and all(l.n_bits_A == 8 for l in lins2)
        assert torch.equal(again.generate(["p"], lat=lat, num_inference_steps=1), out)


def test_fused_inventories_keep_the_work():
    """shapes.*_fused: same FLOPs and the same member Linears as the per-Linear inventories; SD1.5 184 calls -> 100 launches."""
    S = importlib.import_module(PKG + ".shapes")
    for plain, fused in ((S.sd15_unet_linears, S.sd15_unet_linears_fused), (S.sdxl_unet_linears, S.sdxl_unet_linears_fused),
                         (S.sd35_mmdit_linears, S.sd35_mmdit_linears_fused)):
        a, b = plain(), fused()
        assert S.total_flops(a) == S.total_flops(b)
        assert sum(c for *_, c in a) == sum(e[4] * len(e[5]) for e in b)
        assert all(sum(e[5]) == e[2] for e in b)
        members = sorted((e[1], n, e[3]) for e in b for _ in range(e[4]) for n in e[5])
        assert members == sorted((m, n, k) for _, m, n, k, c in a for _ in range(c))
    sd15 = S.sd15_unet_linears_fused()
    assert sum(e[4] for e in sd15) == 100
    by_name = {e[0]: e for e in sd15}
    assert by_name["attn2.to_kv(all blocks)"][1:4] == (1232, 24960, 768) and len(by_name["attn2.to_kv(all blocks)"][5]) == 32
    assert by_name["time_emb_proj(all resnets)"][1:4] == (16, 17600, 1280)
    assert by_name["C320.attn1.to_qkv"][1:5] == (65536, 960, 320, 5)


@pytest.mark.parametrize("kind", ["sd15", "sdxl", "sd35"])
def test_fuse_layers_host_logic(kind):
    """fuse_layers() on a packed skeleton (oracle-backed ops): the fused tensors are the members' concatenated along N
    (utils/fused_utils.py:87-96), the state dict is unchanged, the denoised latents equal the unfused model's, and
    unfuse_layers() / a new quantize() drop the fused copies."""
    fu = importlib.import_module(PKG + ".fused_utils")
    with patched_ops():
        M, model = tiny_model(kind)
        lat = torch.randn(2, model.pipeline.latent_channels, model.pipeline.latent_size, model.pipeline.latent_size,
                          generator=torch.Generator().manual_seed(6)).half()
        model.quantize(quant_config={"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "gemm"}, quantType="awq")
        ref = model.generate(["a", "b"], lat=lat, num_inference_steps=2)
        keys = set(model.denoiser().state_dict().keys())
        done = model.fuse_layers()
        assert done["self_qkv"] > 0 and (done["adaln"] > 0 if kind == "sd35" else done["context_kv"] > 0 and done["time_emb_proj"] > 0)
        assert set(model.denoiser().state_dict().keys()) == keys
        den = model.denoiser()
        n_checked = 0
        for m in den.modules():
            qkv = m.__dict__.get("_qkv")
            if qkv is not None:
                for name in ("qweight", "qzeros", "scales"):
                    assert torch.equal(getattr(qkv, name), torch.cat([getattr(l, name) for l in (m.to_q, m.to_k, m.to_v)], dim=1))
                assert qkv.out_features == 3 * m.to_q.out_features
                n_checked += 1
        assert n_checked * 3 <= done["self_qkv"] and n_checked > 0
        out = model.generate(["a", "b"], lat=lat, num_inference_steps=2, fuse_layers=True)
        assert ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item() <= 2e-3
        assert not any("_pre" in m.__dict__ for m in den.modules())          # every grouped slice was consumed by its member
        model.unfuse_layers()
        assert not any(k in m.__dict__ for m in den.modules() for k in ("_qkv", "_add_qkv", "_ctx_kv", "_temb_all", "_adaln_all", "_fused_act"))
        assert torch.equal(model.generate(["a", "b"], lat=lat, num_inference_steps=2), ref)
    with pytest.raises(TypeError):
        fu.fuse_linears([torch.nn.Linear(8, 8)])


def test_fuse_layers_w8a8_host_logic():
    """SmoothQuant + W8A8 modules (oracle-backed ops): q / k / v share the SmoothQuant divisor, so they fuse; the fused model's
    latents equal the unfused model's."""
    with patched_ops():
        M, model = tiny_sd15()
        lat = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(6)).half()
        model.calib_samples = model.default_calib_samples(1, 2)
        model.quantize(quant_config={"w_bit": 8, "version": "w8a8"}, quantType="sq", alpha=0.5, calib_num_infer_steps=1)
        ref = model.generate(["a", "b"], lat=lat, num_inference_steps=2)
        done = model.fuse_layers()
        assert done["self_qkv"] > 0 and done["context_kv"] > 0
        out = model.generate(["a", "b"], lat=lat, num_inference_steps=2, fuse_layers=True)
        assert ((out.float() - ref.float()).abs().max() / ref.float().abs().max()).item() <= 2e-3


def test_ratio_scales_all_rows_equal_the_per_point_vectors():
    """AwqQuantizer._ratio_scales_all: the batched scale vectors of the 20-point grid are bit-identical to the per-point
    `_ratio_scales` (quantizer.py:717-725), incl. zero / inf statistics, duo and single scaling, every dtype."""
    Q = importlib.import_module(PKG + ".quantizer")
    for duo in (True, False):
        q = Q.AwqQuantizer.__new__(Q.AwqQuantizer)
        q.duo_scaling = duo
        for dt in (torch.float32, torch.float16, torch.bfloat16):
            g = torch.Generator().manual_seed(0)
            xm, wm = (torch.rand(640, generator=g) * 3).to(dt), torch.rand(640, generator=g).to(dt)
            xm[5], wm[7], xm[9] = 0, 0, float("inf")
            rows = q._ratio_scales_all(xm, wm, list(range(20)))
            assert all(torch.equal(rows[i], q._ratio_scales(xm, wm, i / 20)) for i in range(20))
            sub = q._ratio_scales_all(xm, wm, [3, 7, 19])
            assert all(torch.equal(sub[j], q._ratio_scales(xm, wm, i / 20)) for j, i in enumerate([3, 7, 19]))
