"""Multi-GPU: one process per GPU (`torch.distributed`, NCCL over NVLink; gloo in the CPU tests).

The path shards by independent units: the input of block i+1 is the output of the
UN-quantised block i and every group search starts from the original weights, so (block, scaling-group)
searches have no data dependence.  Blocks are assigned to ranks by a cost-balanced greedy partition; each rank
searches its blocks locally (all reductions of one group stay on one GPU, so results are identical for every
world size) and ONE all_gather per phase exchanges {best_scales[K], clip_max[N, G]} -- the only collective.
Prompt-batched denoising is data parallel with no collective in the step loop.
"""
import torch
import torch.distributed as dist


def assign_blocks(names, costs, world):
    """Longest-processing-time greedy partition; deterministic (ties broken by block order) so every rank
    derives the same assignment without communicating."""
    order = sorted(range(len(names)), key=lambda i: (-costs[i], i))
    loads, parts = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        parts[r].append(i)
        loads[r] += costs[i]
    return [[names[i] for i in sorted(p)] for p in parts]


def split_ratios(n_grid, world):
    """ratio-grid split for when there are fewer blocks than ranks: contiguous chunks (20 -> 3,3,3,3,2,2,2,2)."""
    base, extra = divmod(n_grid, world)
    out, s = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(list(range(s, s + n)))
        s += n
    return out


def _flatten(results):
    """{block: {"scales": [(prev, layers, t)], "clip": [(name, t)]}} -> (picklable meta, flat fp32 payload)"""
    meta, chunks = [], []
    for bname, res in results.items():
        sc = [(prev, tuple(layers), tuple(t.shape), str(t.dtype)) for prev, layers, t in res["scales"]]
        cl = [(name, tuple(t.shape), str(t.dtype)) for name, t in res["clip"]]
        meta.append((bname, sc, cl))
        chunks += [t.detach().reshape(-1).float() for _, _, t in res["scales"]]
        chunks += [t.detach().reshape(-1).float() for _, t in res["clip"]]
    return meta, chunks


def _unflatten(meta, payload, device):
    out, off = {}, 0
    for bname, sc, cl in meta:
        scales, clips = [], []
        for prev, layers, shape, dt in sc:
            n = 1
            for s in shape:
                n *= s
            scales.append((prev, layers, payload[off:off + n].to(getattr(torch, dt.split(".")[-1])).reshape(shape).to(device)))
            off += n
        for name, shape, dt in cl:
            n = 1
            for s in shape:
                n *= s
            clips.append((name, payload[off:off + n].to(getattr(torch, dt.split(".")[-1])).reshape(shape).to(device)))
            off += n
        out[bname] = {"scales": scales, "clip": clips}
    return out


def gather_results(local, device):
    """all ranks end with the union of every rank's search results.  Payload: one padded flat fp32 tensor per
    rank through dist.all_gather (fp16/bf16 -> fp32 -> back is exact); metadata through all_gather_object."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    meta, chunks = _flatten(local)
    flat = torch.cat(chunks) if chunks else torch.zeros(0)
    comm_dev = device if dist.get_backend() == "nccl" else torch.device("cpu")
    flat = flat.to(comm_dev)
    metas = [None] * world
    dist.all_gather_object(metas, (meta, flat.numel()))
    maxn = max(n for _, n in metas)
    pad = torch.zeros(maxn, dtype=torch.float32, device=comm_dev)
    pad[:flat.numel()] = flat
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    merged = {}
    for (m, n), buf in zip(metas, bufs):
        merged.update(_unflatten(m, buf[:n], device))
    return merged


def sharded_search(quantizer, shard):
    """run quantizer.search on this rank's blocks and gather; every rank returns the full, identical result."""
    local = quantizer.search(shard=shard)
    dev = next(quantizer.awq_model.denoiser().parameters()).device
    merged = gather_results(local, dev)
    # deterministic application order = model order
    order = list(quantizer.awq_model.get_search_blocks())
    return {k: merged[k] for k in order if k in merged}


def allreduce_hook_stats(hook_dict):
    """Data-parallel SmoothQuant calibration: every rank runs the FP model on its own share of the
    calibration prompts with `calib_data.Fused_Mean_Max_Activation_Hook`s attached; this sums the hooks' fp64
    accumulators (sum over calls of the per-call column max, |x| sum) and counters over the ranks and maxes the running
    max, in place, so that every rank ends with the statistic of the whole calibration set.  The fp64 sum of fp16 maxima
    is exact, hence independent of the reduction order NCCL picks: the smoothing scales -- and the quantised codes
    -- are bit-identical for every world size.  `hook_dict`: {block: {linear: hook}} or {linear: hook}; one
    all_reduce per tensor kind (flattened), not per hook."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return hook_dict
    hooks = []
    for v in hook_dict.values():
        hooks += list(v.values()) if isinstance(v, dict) else [v]
    hooks = [h for h in hooks if getattr(h, "acc_maxsum", None) is not None]
    if not hooks:
        return hook_dict
    dev = hooks[0].acc_maxsum.device
    comm_dev = dev if dist.get_backend() == "nccl" else torch.device("cpu")
    sums = [h.acc_maxsum for h in hooks] + [h.acc_abssum for h in hooks if h.acc_abssum is not None]
    flat = torch.cat([t.reshape(-1) for t in sums]).to(comm_dev)
    counts = torch.tensor([[h.step, h.rows] for h in hooks], dtype=torch.int64, device=comm_dev).reshape(-1)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    off = 0
    for t in sums:
        t.copy_(flat[off:off + t.numel()].reshape(t.shape).to(t.device))
        off += t.numel()
    counts = counts.reshape(-1, 2).tolist()
    for h, (step, rows) in zip(hooks, counts):
        h.step, h.rows = int(step), int(rows)
    # running max: max is exact in any order; reduce in fp32 (fp16 / bf16 -> fp32 -> back is exact)
    rmax = torch.cat([h.running_max.reshape(-1).float() for h in hooks]).to(comm_dev)
    dist.all_reduce(rmax, op=dist.ReduceOp.MAX)
    off = 0
    for h in hooks:
        n = h.running_max.numel()
        h.running_max.copy_(rmax[off:off + n].to(h.running_max.device, h.running_max.dtype))
        off += n
    return hook_dict
