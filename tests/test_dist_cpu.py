"""CPU, world_size 2, gloo: the sharding logic of the calibration search (block assignment, result gather)."""
import importlib
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "quantization---diffusion-models_b200"


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = importlib.import_module(PKG + ".dist")
    names = [f"blk{i}" for i in range(5)]
    mine = d.assign_blocks(names, [5, 1, 4, 2, 3], world)[rank]
    local = {}
    for n in mine:
        i = int(n[3:])
        local[n] = {"scales": [("norm1", ("attn1.to_q", "attn1.to_k"), torch.full((8,), float(i), dtype=torch.float16)),
                               (("adaln", "norm1.linear", (0, 8), (8, 16)), ("ff.net.0.proj",), torch.arange(8, dtype=torch.bfloat16) + i)],
                    "clip": [("ff.net.2", torch.full((4, 2, 1), 0.5 + i, dtype=torch.float16))]}
    merged = d.gather_results(local, torch.device("cpu"))
    ok = sorted(merged) == names
    for n in names:
        i = int(n[3:])
        s0 = merged[n]["scales"][0]
        ok &= s0[0] == "norm1" and tuple(s0[1]) == ("attn1.to_q", "attn1.to_k")
        ok &= s0[2].dtype == torch.float16 and bool((s0[2] == i).all())
        s1 = merged[n]["scales"][1]
        ok &= tuple(s1[0]) == ("adaln", "norm1.linear", (0, 8), (8, 16)) and s1[2].dtype == torch.bfloat16
        ok &= bool(torch.equal(s1[2], torch.arange(8, dtype=torch.bfloat16) + i))
        c = merged[n]["clip"][0]
        ok &= c[0] == "ff.net.2" and tuple(c[1].shape) == (4, 2, 1) and bool((c[1] == 0.5 + i).all())
    q.put((rank, mine, bool(ok)))
    dist.destroy_process_group()


def test_gather_results_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, _, ok in res)
    a, b = res[0][1], res[1][1]
    assert sorted(a + b) == [f"blk{i}" for i in range(5)] and not set(a) & set(b)


def test_assign_blocks_balanced_and_deterministic():
    d = importlib.import_module(PKG + ".dist")
    names = [f"b{i}" for i in range(38)]
    costs = [100] * 37 + [60]
    parts = d.assign_blocks(names, costs, 8)
    assert sorted(sum(parts, [])) == sorted(names)
    loads = [sum(costs[names.index(n)] for n in p) for p in parts]
    assert max(loads) - min(loads) <= 100
    assert parts == d.assign_blocks(names, costs, 8)
    assert d.split_ratios(20, 8) == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11], [12, 13], [14, 15], [16, 17], [18, 19]]
    assert d.assign_blocks(["a"], [1], 1) == [["a"]]


def test_host_logic_cpu():
    """module helpers, config defaults, shape inventories (no device needed)."""
    cfg = importlib.import_module(PKG + ".config").AwqConfig
    c = cfg.from_dict({"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "GEMM"})
    assert c.version == "gemm" and c.a_bit == 16 and c.weight_quant_type == "group" and c.weight_quant_conv_type == "per_channel"
    assert cfg.from_dict({}).version == "fake_act"
    with pytest.raises(TypeError):
        cfg.from_dict({"bogus": 1})
    sh = importlib.import_module(PKG + ".shapes")
    assert abs(sh.total_flops(sh.sd15_unet_linears()) / 1e12 - 3.73) < 0.01
    assert abs(sh.total_flops(sh.sdxl_unet_linears()) / 1e12 - 34.8) < 0.1
    assert abs(sh.total_flops(sh.sd35_mmdit_linears()) / 1e12 - 23.9) < 0.1
    assert sh.group_for(320) == 64 and sh.group_for(2432) == 128
    mod = importlib.import_module(PKG + ".module")
    net = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.Sequential(torch.nn.Conv2d(1, 1, 1), torch.nn.Linear(4, 2)))
    assert list(mod.get_named_linears(net)) == ["0", "1.1"]
    assert mod.get_op_name(net, net[1][1]) == "1.1" and mod.get_op_by_name(net, "1.0") is net[1][0]
    mod.set_op_by_name(net, "1.1", torch.nn.Identity())
    assert isinstance(net[1][1], torch.nn.Identity)
    with pytest.raises(ValueError):
        mod.get_op_by_name(net, "nope")
    assert [n for _, n, _ in mod.get_lin_conv_layers("root", net, None)] == ["0", "0"]
    sk = importlib.import_module(PKG + ".skeletons")
    with torch.device("meta"):
        u = sk.sd15_unet()
    lin = [m for m in u.modules() if isinstance(m, torch.nn.Linear)]
    assert len(lin) == 184 and abs(sum(m.weight.numel() for m in lin) / 1e6 - 270.0) < 0.1
