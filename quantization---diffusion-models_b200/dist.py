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
    """run quantizer.search on this rank's share (blocks, or ratios when there are fewer than 2 blocks per rank) and
    gather; every rank returns the full, identical result."""
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


# ------------------------------------------------------------------ data-parallel capture of the calibration inputs
def owners_of(names, costs, world):
    """(parts, ratio_split): the block -> rank assignment of the search.  With fewer than 2 blocks per rank the ratio grid
    is split instead: every rank searches every block on its share of the 20 ratios."""
    ratio_split = world > 1 and len(names) < 2 * world
    return assign_blocks(names, costs, world), ratio_split


def exchange_captures(caps, wanted_by, rank, world, device):
    """Calibration batches are run data parallel (batch i on rank i % world); this hands every rank the captured Linear
    inputs of the blocks it searches, from ALL batches.
      caps      : {block: {linear: [(call_id, X[rows, K])]}} captured on this rank (call_id = global forward-call index)
      wanted_by : {block: [ranks]} -- who needs the block's inputs (one owner, or every rank when the ratio grid is split)
    Returns the same structure for the blocks `rank` wants, holding the chunks of every rank sorted by call_id, so the
    concatenation is the one a single process would have built: search results do not depend on the world size.
    One metadata all_gather_object + one all_gather per block (NCCL; gloo in the CPU tests)."""
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return {b: {ln: sorted(ch, key=lambda t: t[0]) for ln, ch in lins.items()} for b, lins in caps.items() if rank in wanted_by.get(b, ())}
    meta = {b: {ln: [(cid, tuple(x.shape), str(x.dtype)) for cid, x in ch] for ln, ch in lins.items()} for b, lins in caps.items()}
    metas = [None] * world
    dist.all_gather_object(metas, meta)
    nccl = dist.get_backend() == "nccl"
    comm_dev = device if nccl else torch.device("cpu")
    # One all_gather per block over the communicator's collective channels (NVLS / ring over NVSwitch), not point-to-point
    # sends: the first send/recv between two ranks makes NCCL open a new peer connection -- measured 1.7 s at world 4 and
    # ~8 s at world 8 for 56 pairs, against ~40 ms to move the ~10 GB of SD3.5-L captures through all_gather -- and in
    # ratio-split mode every rank wants every block anyway.  Memory: ONE send and ONE receive buffer (largest block) reused
    # for every block, ONE output arena for everything this rank keeps; chunks are copied into the arena in call order,
    # so the later merge is a view (fresh per-Linear allocations cost ~1 s of cudaMalloc per 14 GB).
    order = sorted(caps)   # identical module trees -> identical keys and order on every rank
    sizes = {b: [sum(sh[0] * sh[1] for ch in metas[r][b].values() for _, sh, _ in ch) for r in range(world)] for b in order}
    dts = {d for r in range(world) for b in order for ch in metas[r][b].values() for _, _, d in ch}
    if not dts:
        return {b: {} for b in order if rank in wanted_by.get(b, ())}
    assert len(dts) == 1, f"captures have mixed dtypes {dts}"
    dt = getattr(torch, dts.pop().split(".")[-1])
    n_cap = max(max(v) for v in sizes.values())
    send = torch.zeros(n_cap, dtype=dt, device=comm_dev)
    recv = torch.empty(world * n_cap, dtype=dt, device=comm_dev) if nccl else None
    arena = torch.empty(sum(sum(sizes[b]) for b in order if rank in wanted_by.get(b, ())), dtype=dt, device=device)
    pos = 0
    mine = {}
    for b in order:
        n_max = max(sizes[b])
        if n_max == 0:
            continue
        parts = [x.reshape(-1) for ch in caps[b].values() for _, x in ch]
        if parts:
            flat = torch.cat(parts)
            send[: flat.numel()].copy_(flat)
        if nccl:
            dist.all_gather_into_tensor(recv[: world * n_max], send[:n_max])
            bufs = [recv[r * n_max:(r + 1) * n_max] for r in range(world)]
        else:
            bufs = [torch.empty(n_max, dtype=dt) for _ in range(world)]
            dist.all_gather(bufs, send[:n_max].clone())
        if rank not in wanted_by.get(b, ()):
            continue
        entries = {}          # linear -> [(call id, peer, offset in the peer's buffer, shape)]
        for peer in range(world):
            off = 0
            for ln, ch in metas[peer][b].items():
                for cid, shape, _ in ch:
                    entries.setdefault(ln, []).append((cid, peer, off, shape))
                    off += shape[0] * shape[1]
        mine[b] = {}
        for ln in metas[rank][b] if metas[rank][b] else entries:
            chunks = []
            for cid, peer, off, shape in sorted(entries.get(ln, ())):
                n = shape[0] * shape[1]
                dst = arena[pos:pos + n].view(shape)
                dst.copy_(bufs[peer][off:off + n].view(shape))
                pos += n
                chunks.append((cid, dst))
            mine[b][ln] = chunks
    return mine


def allreduce_min_losses(losses):
    """ratio-split search: every rank filled its own ratios of the [groups, 20] loss table (inf elsewhere); ONE
    all_reduce(MIN) completes it everywhere (the entries are disjoint, so MIN just merges them, bit for bit)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return losses
    if dist.get_backend() == "nccl":
        dist.all_reduce(losses, op=dist.ReduceOp.MIN)
        return losses
    t = losses.cpu()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.to(losses.device)
