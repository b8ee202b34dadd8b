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
        fp = model.generate(["a", "b"], lat=lat, num_inference_steps=2).float()
        n_pw = sum(1 for m in model.denoiser().modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (1, 1))
        n_33 = sum(1 for m in model.denoiser().modules() if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3))
        model.quantize(quant_config={"q_group_size": 64, "w_bit": 4 if version == "gemm" else 8, "version": version},
                       quantType="awq")
        mods = [m for m in model.denoiser().modules() if type(m).__name__ == "QConv1x1"]
        assert len(mods) == n_pw > 0 and all(type(m.inner).__name__ == inner for m in mods)
        # 3x3 / stride 1 / pad 1 convolutions with C % 64 == 0 carry packed int4 weights too (implicit-GEMM kernel c);
        # conv_in (C = 4), conv_out and the stride-2 down-samplers keep fake-quant weights
        n_q3 = sum(1 for m in model.denoiser().modules() if type(m).__name__ == "QConv3x3")
        n_fake = sum(1 for m in model.denoiser().modules() if type(m).__name__ == "WxAxConv2d")
        assert n_q3 + n_fake == n_33 and n_fake >= 2
        assert (n_q3 > 0) == (version == "gemm")
        assert not any(isinstance(m, (torch.nn.Linear, torch.nn.Conv2d)) for m in model.denoiser().modules())
        out = model.generate(["a", "b"], lat=lat, num_inference_steps=2)
        assert torch.isfinite(out).all()
        assert ((out.float() - fp).abs().max() / fp.abs().max()).item() < (0.5 if version == "gemm" else 0.1)
        model.save_quantized(str(tmp_path))
        again = M.StableDiffusion1_x.from_quantized(str(tmp_path), device="cpu")
        for k, v in packed_state(model).items():
            assert torch.equal(v, again.denoiser().state_dict()[k]), k
        assert torch.equal(again.generate(["a", "b"], lat=lat, num_inference_steps=2), out)


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


def _calibrated(shard):
    M, model = tiny_sd15()
    model.calib_steps = 2
    model.quantize(quant_config={"zero_point": True, "q_group_size": 64, "w_bit": 4, "version": "gemm"}, quantType="awq",
                   calibrate=True, shard=shard)
    return model


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    with patched_ops():
        model = _calibrated((rank, world))
    q.put((rank, {k: v.numpy() for k, v in packed_state(model).items()}, len(model.quantizer.search_log)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_awq_search_world2_equals_world1():
    """blocks are searched on the rank that owns them, ONE gather exchanges {scales, clip}, every rank
    applies the identical list -> the packed codes / zeros / scales are bit-identical for world 1 and world 2, on
    both ranks, and each rank searched only its own blocks."""
    with patched_ops():
        single = _calibrated(None)
    want = packed_state(single)
    n_groups = len(single.quantizer.search_log)
    assert n_groups == 3 * len(single.get_search_blocks())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
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
