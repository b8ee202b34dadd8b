#!/usr/bin/env python
"""bench.py -- the quantized-linear hot path on B200.

Workload (BASELINE.json configs[1]): every nn.Linear of the SD1.5 UNet at 512x512, batch 8 with CFG
(B_eff 16), W4A16 AWQ group 128 (64 where K=320, quantize/fake_quant.py:34-37).  One "step" = one pass
over all 184 Linear calls of one denoise step (43 distinct shapes, 3.73 TFLOP), each through
WQLinear_GEMM.forward -> qdm_gemm_w4a16 (tcgen05 dequant-in-mainloop GEMM).  Synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sweep] [--layers]

  value     whole-job TFLOP/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       same pass through the module API with the step's inputs staged from pinned HOST memory
            and the step's final output read back to the host inside the timed region
  roofline  W4A16 GEMM kernel: algorithmic 2MNK flops / CUDA-event time vs measured cuBLAS bf16 peak
  cpu_baseline / --impl reference: the reference's CPU torch path (F.linear on fake-quant weights,
            quantize/fake_quant.py:223) timed on the host cores on a bounded sample of the same layers
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "w4a16_qlinear_tflops"
UNIT = "TFLOP/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"], "hbm": p["hbm_gbs"],
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def step_traffic(n_launches=None):
    """DRAM bytes of one step (sum over its launches) from the newest committed ncu launch list whose launch count is the
    step's, or None."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "step_traffic_r*.json")), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
            if n_launches is None or d.get("launches", sum(d.get("kernels", {}).values())) == n_launches:
                return d["dram_bytes_per_step"]
        except (OSError, KeyError, ValueError):
            continue
    return None


def layer_list(fused=False):
    """The step's Linear calls.  fused=True: the same Linears with the same-input ones (attn1 q/k/v per block, attn2 k/v of all
    blocks, the 22 time_emb_proj) as one launch each on the N-concatenated packed weight (fused_utils.fuse_linears; the
    reference's utils/fused_utils.py:87-96): entries carry a 6th field, the members' N."""
    shapes = importlib.import_module("quantization---diffusion-models_b200.shapes")
    return shapes, (shapes.sd15_unet_linears_fused(batch=8, cfg=True) if fused else shapes.sd15_unet_linears(batch=8, cfg=True))


def per_shape_roofline_ms(shapes, layers, peaks):
    """sum over the step's launches of max(flops / tensor peak, algorithmic bytes / HBM peak), in ms"""
    tot = 0.0
    for _, m, n, k, c in (e[:5] for e in layers):
        by = shapes.gemm_bytes_w4a16(m, n, k, shapes.group_for(k))
        tot += c * max(2.0 * m * n * k / (peaks["bf16_burst"] * 1e12), by / (peaks["hbm"] * 1e9))
    return tot * 1e3


def workload_config(shapes, layers, world):
    """`config` of the JSON line -- identical for the GPU arm and the reference arm (the reference arm times a bounded
    sample of this workload; what the sample was is said in its `cpu_baseline.sample`)."""
    return {"workload": "sd15_unet_w4a16_linears_b8cfg",
            "reference_config": "SD1.5 UNet W4A16 AWQ group-128, 512x512 latents batch 8 (CFG -> B_eff 16)",
            "linear_calls_per_step": sum(e[4] for e in layers), "distinct_shapes": len(layers),
            "tflop_per_step": shapes.total_flops(layers) / 1e12, "group_size": "128 (64 where K=320)",
            "l2": "per-step working set (>1.5 GB of activations + 184 distinct weights) exceeds the 126 MB L2",
            "parallelism": f"dp{world} (prompt-batched, no collective in the step)"}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(sample_budget_s=20.0, steps=1, warmup=0, full=False):
    """The reference's CPU torch path for this workload: F.linear(x, W_fakequant, bias) in fp32 math on all
    host cores (WxAxLinear.forward, quantize/fake_quant.py:223; the oracle restates it as linear_fake).
    full=False (the `cpu_baseline` of the GPU arm's line): a bounded sample, one call per distinct layer shape with M
    capped at 4096 rows.  full=True (`--impl reference`): the WHOLE step -- every one of the step's 184 Linear calls at its
    full M (the calls of one distinct shape reuse that shape's tensors; the pass still streams > 1 GB of activations) --
    so that the reference arm's config is the GPU arm's; the time budget only limits how many timed passes run."""
    import torch
    import oracle.qdm_oracle as O

    shapes, layers = layer_list()
    torch.set_num_threads(os.cpu_count() or 1)
    cap_m = None if full else 4096
    sample = [(n, m if full else min(m, cap_m), nn_, k, c if full else 1) for n, m, nn_, k, c in layers]
    flops = sum(2.0 * m * nn_ * k * c for _, m, nn_, k, c in sample)
    g = torch.Generator().manual_seed(42)
    data, xs = [], {}
    for name, m, nn_, k, c in sample:
        w = (torch.randn(nn_, k, generator=g) * 0.02).half()
        wq = O.rtn_group(w, shapes.group_for(k), True, 4)[0]
        if (m, k) not in xs:
            xs[(m, k)] = torch.randn(m, k, generator=g).half()
        data.append((xs[(m, k)], wq, torch.zeros(nn_).half(), c))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for x, wq, b, c in data:
            for _ in range(c):
                O.linear_fake(x, wq, b)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        if sum(times) > sample_budget_s:
            break
    dt = sum(times) / len(times)
    what = (f"the whole step: all {sum(c for *_, c in sample)} F.linear calls of the SD1.5 UNet step at full M "
            f"({len(sample)} distinct shapes, each shape's calls on that shape's tensors)" if full else
            f"one F.linear per distinct SD1.5 UNet Linear shape ({len(sample)} shapes), M capped at {cap_m} rows")
    return {"value": flops / dt / 1e12, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{what}, {flops / 1e9:.1f} GFLOP per pass, fp16 fake-quant weights, fp32 math",
            "ms_per_pass": dt * 1e3, "passes": len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shapes, layers = layer_list()
    cb = cpu_reference(sample_budget_s=150.0, steps=max(1, args.steps), warmup=max(0, args.warmup), full=True)
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["passes"],
            "warmup": args.warmup, "ms_per_step": cb["ms_per_pass"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "impl": "reference",
            "config": workload_config(shapes, layers, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def build_layers(q, shapes, layers, dev, dtype):
    """random-init weights of every Linear of the step (distinct per layer instance, like the real UNet), AWQ-quantised on
    the GPU by the fused RTN+pack kernel; one activation buffer per distinct (M, K); outputs come from the op itself
    (allocated by the API like the reference's F.linear).  `layers` may be the fused inventory (6-field entries): then every
    member Linear is built as its own WQLinear_GEMM and the launch is fused_utils.fuse_linears of the members.
    Returns (xs, launches, members): `launches` = the step as launched, `members` = the same Linears one by one."""
    import torch
    lin = importlib.import_module("quantization---diffusion-models_b200.linear")
    fu = importlib.import_module("quantization---diffusion-models_b200.fused_utils")
    g = torch.Generator(device=dev).manual_seed(42)
    xs, launches, members = {}, [], []
    for e in layers:
        name, m, n, k, cnt = e[:5]
        parts = e[5] if len(e) > 5 else (n,)
        if (m, k) not in xs:
            xs[(m, k)] = torch.randn(m, k, generator=g, device=dev, dtype=dtype)
        for inst in range(cnt):
            mem = []
            for pn in parts:
                fl = torch.nn.Linear(k, pn, bias=True, device=dev, dtype=dtype)
                fl.weight.data = torch.randn(pn, k, generator=g, device=dev, dtype=dtype) * 0.02
                mem.append(lin.WQLinear_GEMM.from_linear(fl, 4, shapes.group_for(k)))
                del fl
            members += [(name, mm, (m, k), inst) for mm in mem]
            launches.append((name, mem[0] if len(mem) == 1 else fu.fuse_linears(mem), (m, k), inst))
    return xs, launches, members



# ------------------------------------------------------------------------------------------ sub-records of the bench line
def _graph_of(torch, fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


def int8_dense_peak(torch, dev):
    """Dense INT8 tensor-pipe peak of this GPU (BASELINE.md section 3 asks for a measured denominator of W8A8): cuBLASLt
    int8 x int8 -> int32 through torch._int_mm on 8192^3, best of 10 (burst) -- the same protocol as MEASURED_PEAKS'
    bf16 number.  None if the library path is unavailable."""
    try:
        n = 8192
        a = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
        b = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev).t().contiguous().t()   # column-major operand
        for _ in range(3):
            torch._int_mm(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch._int_mm(a, b); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / best / 1e9
    except Exception:
        return None


def sub_w8a8(q, shapes, layers, dev, dtype, steps, warmup, timed):
    """Kernel (d) on the same workload: every Linear of the step through W8A8Linear.forward = per-token int8 quantiser
    (qdm_actquant_token_i8) + int8 tcgen05 GEMM with the dequant scales in the epilogue; one CUDA graph.  `layers` is the fused
    inventory: same-input Linears are one W8A8Linear (fused_utils.fuse_linears), so a shared input is quantised once.
    Also the GEMMs alone on pre-quantised activations, the per-Linear (unfused) step, and the measured INT8 dense peak."""
    import torch
    lin = importlib.import_module("quantization---diffusion-models_b200.linear")
    fu = importlib.import_module("quantization---diffusion-models_b200.fused_utils")
    g = torch.Generator(device=dev).manual_seed(43)
    xs, mods, members = {}, [], []
    for e in layers:
        name, m, n, k, cnt = e[:5]
        parts = e[5] if len(e) > 5 else (n,)
        if k % 16:
            continue
        if (m, k) not in xs:
            xs[(m, k)] = torch.randn(m, k, generator=g, device=dev, dtype=dtype)
        for _ in range(cnt):
            mem = []
            for pn in parts:
                fl = torch.nn.Linear(k, pn, bias=True, device=dev, dtype=dtype)
                fl.weight.data = torch.randn(pn, k, generator=g, device=dev, dtype=dtype) * 0.02
                mem.append(lin.W8A8Linear.from_float(fl))
                del fl
            members += [(mm, (m, k)) for mm in mem]
            mods.append((mem[0] if len(mem) == 1 else fu.fuse_linears(mem), (m, k)))
    flops = sum(2.0 * key[0] * mod.out_features * key[1] for mod, key in mods)

    def step_of(lst):
        def step():
            for mod, key in lst:
                mod(xs[key])
        return step

    step = step_of(mods)
    pre = {key: q.ops.actquant_token_i8(x) for key, x in xs.items()}

    def step_gemm_only():
        for mod, key in mods:
            xq, sx = pre[key]
            q.ops.gemm_w8a8(xq, sx, mod.qweight, mod.w_scales, mod.bias, out_dtype=mod.out_dtype)

    q.ops.launch_count(reset=True)
    step()
    launches = q.ops.launch_count()
    ms = timed(_graph_of(torch, step).replay, steps, warmup) / steps
    ms_g = timed(_graph_of(torch, step_gemm_only).replay, steps, warmup) / steps
    ms_u = timed(_graph_of(torch, step_of(members)).replay, steps, warmup) / steps if len(members) != len(mods) else ms
    peak = int8_dense_peak(torch, dev)
    tf, tf_g = flops / ms / 1e9, flops / ms_g / 1e9
    # per-shape roofline of the GEMMs: max(int8 tensor time, algorithmic bytes at the measured HBM rate), summed over the launches
    pk = measured_peaks()
    roof_ms = sum(max(2.0 * key[0] * mod.out_features * key[1] / ((peak or 3000.0) * 1e12),
                      shapes.gemm_bytes_w8a8(key[0], mod.out_features, key[1]) / (pk["hbm"] * 1e9)) for mod, key in mods) * 1e3
    return {"tflops": tf, "ms_per_step": ms, "gemm_only_tflops": tf_g, "gemm_only_ms_per_step": ms_g, "launches_per_step": launches,
            "gemm_only_frac_per_shape_roofline": roof_ms / ms_g,
            "unfused_ms_per_step": ms_u, "unfused_tflops": flops / ms_u / 1e9, "linears_per_step": len(members),
            "tflop_per_step": flops / 1e12, "int8_dense_peak_tops": peak, "peak_how": "torch._int_mm (cuBLASLt int8) 8192^3, best of 10, this run",
            "frac_of_int8_peak": (tf_g / peak) if peak else None,
            "what": "W8A8Linear.forward over the step's Linears (same-input ones fused: one activation quantisation + one GEMM): "
                    "qdm_actquant_token_i8 + qdm_gemm_w8a8 (quantize/fake_quant.py:86-118); unfused_* = every Linear on its own"}


def sub_config5(q, shapes, dev):
    """BASELINE config 5 in the driver's own line: four corners of the GEMM sweep (`bench.py --sweep` has all 33 shapes,
    profiles/gemm_sweep_r02.json), W4A16 (kernel c, module dispatch) and the W8A8 GEMM (kernel d) against cuBLAS bf16 on the same
    shapes.  GPU-side time of one launch with a cold L2: CUDA graph of 10 x (256 MB flush, launch) minus the flushes alone.
    Rank-local (no collective); never raises."""
    import torch
    try:
        peaks = measured_peaks()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def graph_ms(body, reps=10):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                for _ in range(reps):
                    body()
            g_.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g_.replay(); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            return best

        flush_ms = graph_ms(lambda: flush.zero_())

        def time_fn(fn):
            def body():
                flush.zero_()
                fn()
            return max(graph_ms(body) - flush_ms, 1e-4)

        i8_peak = int8_dense_peak(torch, dev)
        g = torch.Generator(device=dev).manual_seed(45)
        rows = []
        for m, n, k in ((4096, 4096, 4096), (16384, 4096, 4096), (65536, 3072, 3072), (16384, 6144, 1536)):
            grp = shapes.group_for(k)
            x = torch.randn(m, k, generator=g, device=dev, dtype=torch.float16)
            w = torch.randn(n, k, generator=g, device=dev, dtype=torch.float16) * 0.02
            qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
            bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp)
            flops = 2.0 * m * n * k
            t4 = time_fn(lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, None, bts))
            kern = list(q.ops.gemm_last_variant())
            xb, wb = x.bfloat16(), dq.bfloat16()
            tb = time_fn(lambda: torch.nn.functional.linear(xb, wb))
            del xb, wb
            xq, sx = q.ops.actquant_token_i8(x)
            _, wq, sw, _ = q.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
            swf = sw.float()
            t8 = time_fn(lambda: q.ops.gemm_w8a8(xq, sx, wq, swf))
            r = {"M": m, "N": n, "K": k, "w4a16_kernel": kern,
                 "w4a16_tflops": flops / t4 / 1e9, "w4a16_frac_bf16_burst_peak": flops / t4 / 1e9 / peaks["bf16_burst"],
                 "cublas_bf16_tflops": flops / tb / 1e9, "w8a8_gemm_tops": flops / t8 / 1e9,
                 "w8a8_frac_int8_peak": (flops / t8 / 1e9 / i8_peak) if i8_peak else None}
            rows.append(r)
            del x, w, qw, qz, sc, dq, bts, xq, sx, wq, sw, swf
            torch.cuda.empty_cache()
        return {"rows": rows, "bf16_burst_peak_tflops": peaks["bf16_burst"], "int8_dense_peak_tops": i8_peak,
                "peak_how": "bf16: MEASURED_PEAKS.json burst (fallback if absent); int8: torch._int_mm 8192^3 best of 10, this run",
                "what": "BASELINE config 5 corners: one launch, cold L2, GPU-side graph timing; the full 33-shape table is bench.py --sweep"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def sub_denoise(args, dev, rank, world, sync_max):
    """BASELINE config 2: SD1.5 UNet skeleton, W4A16 AWQ g128, 512^2 latents batch 8 + CFG, 50-step loop through generate();
    prompt-batched data parallel, no collective in the loop.  it/s = denoise steps per second per replica."""
    import torch
    M = importlib.import_module("quantization---diffusion-models_b200.models")
    q = importlib.import_module("quantization---diffusion-models_b200")
    model = M.StableDiffusion1_x.from_skeleton(device=dev)
    model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq")
    batch, steps = 8, 50
    prompts = [f"prompt {rank}-{i}" for i in range(batch)]
    lat = torch.randn(batch, 4, 64, 64, generator=torch.Generator().manual_seed(42 + rank)).to(dev, model.pipeline.dtype)
    model.generate(prompts, lat=lat, num_inference_steps=2)
    sync_max(0.0)

    def timed_loop(graph, fuse=False):
        model.generate(prompts, lat=lat, num_inference_steps=2, cuda_graph=graph, fuse_layers=fuse)   # warm-up (captures the step when graph)
        sync_max(0.0)
        q.ops.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = model.generate(prompts, lat=lat, num_inference_steps=steps, cuda_graph=graph, fuse_layers=fuse)
        e1.record()
        torch.cuda.synchronize()
        return sync_max(e0.elapsed_time(e1) * 1e-3), r, q.ops.launch_count()

    def rel(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6))

    sec_eager, res_eager, launches = timed_loop(False)
    sec_graph, res_graph, _ = timed_loop(True)
    fused_members = model.fuse_layers()
    sec_feager, res_feager, launches_fused = timed_loop(False, True)
    sec, res, _ = timed_loop(True, True)
    out = {"it_per_s": steps / sec, "images_it_per_s": world * batch * steps / sec, "steps": steps, "batch_per_gpu": batch, "quant": "w4a16 g128",
           "seconds": sec, "finite": bool(torch.isfinite(res).all()), "it_per_s_eager": steps / sec_eager,
           "it_per_s_graph_unfused": steps / sec_graph, "it_per_s_eager_fused": steps / sec_feager,
           "graph_vs_eager_max_rel": rel(res_graph, res_eager), "fused_vs_unfused_max_rel": rel(res, res_graph),
           "libqdm_launches_per_step": launches_fused // steps, "libqdm_launches_per_step_unfused": launches // steps,
           "fused_members": fused_members,
           "how": "generate(..., cuda_graph=True, fuse_layers=True): every denoiser call replayed from one CUDA graph (skeletons.SkeletonPipeline."
                  "_graph_step), same-input packed projections as one launch each + qdm_geglu (fused_utils.fuse_projections); it_per_s_eager = "
                  "the unfused loop launched eagerly; *_max_rel = max |a - b| / max |b| of the final latents after the 50 steps",
           "model": "SD1.5 UNet skeleton (random init), packed Linears + 1x1 / 3x3 convolutions", "scaling": "weak (data parallel, no collective in the loop)"}
    del model
    torch.cuda.empty_cache()
    return out


class NumaBinding:
    """Context manager: bind this process to the CPUs of the NUMA node GPU `local` hangs off (sysfs), restore on exit.
    Memory pinned inside the block is allocated on that node (first touch / current mempolicy).  A no-op when the
    topology cannot be read (node == None) or the container's cpuset does not include those CPUs."""

    def __init__(self, local):
        self.node, self.cpus, self.saved = None, None, None
        try:
            import torch
            pr = torch.cuda.get_device_properties(local)
            path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
            node = int(open(path).read().strip())
            if node < 0:
                return
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                self.node, self.cpus = node, cpus
        except Exception:
            pass

    def __enter__(self):
        if self.cpus:
            self.saved = os.sched_getaffinity(0)
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        if self.saved:
            os.sched_setaffinity(0, self.saved)
        return False


def _warm_pipeline(model):
    """one untimed 1-step call of the FP pipeline on the first calibration batch (CUDA module load, cuBLAS / SDPA first-call
    setup): the calibration timer starts warm, like every other timed region of this file"""
    import torch
    prompts, latents = model.calib_samples[0]
    with torch.no_grad():
        model.pipeline(prompt=prompts, latents=latents, num_inference_steps=1, guidance_scale=7.5)
    torch.cuda.synchronize()


def sub_calib(args, dev, rank, world, sync_max, model_name="sd35", blocks=0, calib_batches=8, calib_steps=2):
    """BASELINE config 4: AWQ calibration (capture, 20-point scale search, clip search, quantise + pack) of the SD3.5-Large
    MMDiT skeleton (38 blocks, 8.05 G Linear parameters), sharded over the ranks: data-parallel capture, block-sharded
    search, one gather.  codes_checksum covers qweight / qzeros / scales of every packed Linear: equal across world sizes
    = bit-identical codes."""
    import torch
    M = importlib.import_module("quantization---diffusion-models_b200.models")
    cls = {"sd15": M.StableDiffusion1_x, "sdxl": M.StableDiffusionXL, "sd35": M.StableDiffusion3_5}[model_name]
    arch = {"layers": blocks} if (model_name == "sd35" and blocks) else {}
    t0 = time.perf_counter()
    model = cls.from_skeleton(device=dev, **arch)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    batch = {"sd15": 8, "sdxl": 4, "sd35": 1}[model_name]
    model.calib_samples = model.default_calib_samples(calib_batches, batch)
    model.calib_steps = calib_steps
    _warm_pipeline(model)
    sync_max(0.0)
    t0 = time.perf_counter()
    model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq",
                   calibrate=True, shard=(rank, world) if world > 1 else None)
    torch.cuda.synchronize()
    sec = sync_max(time.perf_counter() - t0)
    chk, n_mod = 0, 0
    for name, mod in model.denoiser().named_modules():
        if type(mod).__name__ == "WQLinear_GEMM":
            n_mod += 1
            for t in (mod.qweight, mod.qzeros, mod.scales.view(torch.int16)):
                chk = (chk * 1000003 + int(t.to(torch.int64).sum().item()) + t.numel()) % (1 << 61)
    tm = dict(getattr(model.quantizer, "timings", {}) or {})
    for k in ("capture_s", "capture_forward_s", "capture_exchange_s", "scale_search_s", "clip_search_s"):
        if k in tm:
            tm[k] = sync_max(tm[k])
    out = {"s_per_model": sec, "model": {"sd35": "SD3.5-Large MMDiT skeleton", "sd15": "SD1.5 UNet skeleton", "sdxl": "SDXL UNet skeleton"}[model_name],
           "blocks": len(model.get_search_blocks()), "packed_modules": n_mod, "codes_checksum": chk, "phases_max_over_ranks": tm,
           "calib_batches": calib_batches, "calib_steps": calib_steps, "calib_batch": batch, "world": world, "build_s": build_s, "warm": "one untimed 1-step FP pipeline call before the timer",
           "collectives": "capture exchange (one all_gather per block) + one all_gather of {scales, clip}" if world > 1 else "none"}
    del model
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    q = importlib.import_module("quantization---diffusion-models_b200")
    q._lib.check(q._lib.load().qdm_device_check(local))
    fused = not args.no_fuse
    shapes, layers_unfused = layer_list()
    layers = layer_list(fused=True)[1] if fused else layers_unfused
    dtype = torch.float16
    flops_step = shapes.total_flops(layers_unfused)
    assert abs(shapes.total_flops(layers) - flops_step) < 1e-6 * flops_step   # fusing changes the launches, not the work
    xs, mods, members = build_layers(q, shapes, layers, dev, dtype)
    n_calls = len(mods)
    torch.cuda.synchronize()

    def step_of(lst):
        def step():
            y = None
            for _, mod, key, _ in lst:
                y = mod(xs[key])
            return y
        return step

    step = step_of(mods)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # ---- device-resident throughput.  The launches of a step are captured once in a CUDA graph (the tensor maps are
    # by-value kernel parameters, so the capture is exact) and replayed: per-launch host work (Python, ctypes,
    # cuTensorMapEncodeTiled) would otherwise bound the short small-M launches.
    step()   # first call of every module builds its kernel-native weight copy (WQLinear_GEMM._repacked): not part of a step
    torch.cuda.synchronize()
    run_step = step
    graph = None
    if not args.no_graph:
        graph = _graph_of(torch, step)
        run_step = graph.replay
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    q.ops.launch_count(reset=True)
    ms = timed(run_step, args.steps, args.warmup)
    launches = n_calls * args.steps   # one libqdm kernel per launch of the step (checked against the library's counter below)
    if graph is None:
        assert q.ops.launch_count() == n_calls * (args.steps + args.warmup), q.ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms / args.steps
    value = world * flops_step / (ms_step * 1e-3) / 1e12

    # ---- the same Linears launched one by one (184 launches: the round-1 / round-2 form of this line), for continuity
    unfused = None
    if fused and not args.no_graph:
        step_members = step_of(members)
        step_members()
        torch.cuda.synchronize()
        g_members = _graph_of(torch, step_members)
        ms_members = timed(g_members.replay, args.steps, args.warmup) / args.steps
        unfused = {"ms_per_step": ms_members, "tflops": world * flops_step / (ms_members * 1e-3) / 1e12, "launches_per_step": len(members),
                   "what": "every Linear as its own launch (no same-input fusion), one CUDA graph"}
        del g_members
    for _, mm, _, _ in members:   # the members' kernel-native copies are not needed again
        mm.__dict__.pop("_rp", None)

    # ---- end to end: the step's EXTERNAL inputs from pinned host memory, final output back to the host.  Every Linear
    # reads the host-supplied activation of its (M, K) except ff.net.2, which reads what it reads in the model: the GEGLU of
    # its own block's ff.net.0.proj output, computed on the device (ops.geglu, one HBM pass, inside the timed region and not
    # counted as FLOPs).  The staging buffers are allocated (and first touched) while this process is bound to the CPUs of
    # the GPU's own NUMA node.
    chain_keys = {key for name, _, key, _ in mods if name.endswith("ff.net.2")} if not args.no_chain else set()
    ext_keys = [k for k in xs if k not in chain_keys]
    numa = NumaBinding(local)
    with numa:
        host_x = {k: torch.empty(xs[k].shape, dtype=dtype).pin_memory().copy_(xs[k]) for k in ext_keys}
        y_probe = step()
        host_y = torch.empty(y_probe.shape, dtype=dtype).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host_x.values())
    d2h = host_y.numel() * host_y.element_size()

    # The inputs are staged on a copy stream in order of first use; the layer that first reads an input waits for that copy
    # only, so H2D overlaps the GEMMs of the inputs that have already landed.  The whole step -- the H2D copies, the
    # WQLinear_GEMM.forward launches, the GEGLUs, the D2H read -- is one two-stream CUDA graph.
    # Launch order of the e2e step: the Linears whose inputs are smallest first (the order of a step's Linears is free in this
    # workload; a chained ff.net.2 stays behind its ff.net.0.proj), so the GEMMs start after the first kilobytes have landed
    # instead of after the largest activation.
    def ext_key(name, key):
        return (key[0], key[1] // 4) if (key in chain_keys and name.endswith("ff.net.2")) else key
    in_bytes = lambda k: k[0] * k[1] * 2
    mods_e2e = sorted(mods, key=lambda e: in_bytes(ext_key(e[0], e[2])))   # stable: ties keep the step's order
    copy_stream = torch.cuda.Stream()
    first_use = []
    for _, _, key, _ in mods_e2e:
        if key not in first_use and key not in chain_keys:
            first_use.append(key)
    n_geglu = sum(1 for name, _, key, _ in mods if name.endswith("ff.net.2") and key in chain_keys)

    last_mod = mods[-1][1]   # the step's result read back to the host: the output of the step's last Linear, as in `step`

    def step_e2e():
        cur = torch.cuda.current_stream()
        copy_stream.wait_stream(cur)
        evs = {}
        with torch.cuda.stream(copy_stream):
            for k in first_use:
                xs[k].copy_(host_x[k], non_blocking=True)
                evs[k] = torch.cuda.Event()
                evs[k].record(copy_stream)
        seen = set()
        ff_out = {}
        y_out = None
        for name, mod, key, inst in mods_e2e:
            if key in chain_keys and name.endswith("ff.net.2"):
                y = mod(q.ops.geglu(ff_out.pop((name[:-len("ff.net.2")], key[0], inst))))
            else:
                if key not in seen:
                    cur.wait_event(evs[key])
                    seen.add(key)
                y = mod(xs[key])
                if chain_keys and name.endswith("ff.net.0.proj"):
                    ff_out[(name[:-len("ff.net.0.proj")], key[0], inst)] = y
            if mod is last_mod:
                y_out = y
        host_y.copy_(y_out, non_blocking=True)
        cur.wait_stream(copy_stream)

    run_e2e = step_e2e
    if not args.no_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step_e2e()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph_e2e = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_e2e):
            step_e2e()
        run_e2e = graph_e2e.replay
    ms_e2e = timed(run_e2e, args.steps, max(1, args.warmup // 2)) / args.steps
    e2e_value = world * flops_step / (ms_e2e * 1e-3) / 1e12

    # ---- the other BASELINE metrics as sub-records of the same line (every rank takes part: they shard / run data parallel)
    def sync_max(x):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    variants = {}
    for _, mod, key, _ in mods:   # which kernel each launch of the step runs (dispatch is shape dependent)
        mod(xs[key])
        v, tile = q.ops.gemm_last_variant()
        variants[v] = variants.get(v, 0) + 1
    del graph, mods, members
    if not args.no_graph:
        del graph_e2e
    torch.cuda.empty_cache()
    extra = {}
    if not args.no_extras:
        wanted = [w for w in args.extras.split(",") if w]
        for name, fn in (("w8a8", lambda: sub_w8a8(q, shapes, layers, dev, dtype, args.steps, args.warmup, timed)),
                         ("config5", lambda: sub_config5(q, shapes, dev)),
                         ("denoise", lambda: sub_denoise(args, dev, rank, world, sync_max)),
                         ("calib", lambda: sub_calib(args, dev, rank, world, sync_max, blocks=args.blocks))):
            if name not in wanted:
                continue
            try:
                extra[name] = fn()
            except Exception as e:   # a sub-record must never take the headline metric down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                if world > 1:
                    raise
            torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    achieved = flops_step / (ms_step * 1e-3) / 1e12  # per GPU; every launch of the step is one kernel of the W4A16 family
    alg_bytes = sum(e[4] * shapes.gemm_bytes_w4a16(e[1], e[2], e[3], shapes.group_for(e[3])) for e in layers)
    how = (f"{n_calls} launches: the step's 184 Linears with the same-input ones as one launch each on the N-concatenated packed weight "
           "(fused_utils.fuse_linears, utils/fused_utils.py:87-96: attn1 q/k/v per block, attn2 k/v of all 16 blocks, the 22 time_emb_proj)"
           if fused else "184 launches, one per Linear")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": workload_config(shapes, layers_unfused, world),
        "launch": ("eager, " if args.no_graph else "one CUDA graph, ") + how,
        "launches_per_step": n_calls,
        "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                "host_numa_node": numa.node, "geglu_launches_per_step": n_geglu,
                "inputs": ("every Linear's input from pinned host memory except ff.net.2, which reads geglu(ff.net.0.proj output) of its own "
                           "block computed on the device (qdm_geglu, inside the timed region, not counted as FLOPs)" if chain_keys else
                           "every Linear's input from pinned host memory")},
        # the timed region is short (0.1 s at full clocks, far below the power cap), so the honest denominator is the BURST
        # cuBLAS bf16 peak; frac_sustained and the per-shape roofline (HBM-bound shapes counted at the HBM rate) beside it
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_burst"], "traffic": step_traffic(n_calls),
                     "frac_sustained_peak": achieved / peaks["bf16_sustained"],
                     "frac_per_shape_roofline": per_shape_roofline_ms(shapes, layers, peaks) / ms_step,
                     "algorithmic_bytes_per_step": alg_bytes,
                     "kernel": "W4A16 family, launches per step by kernel: " + json.dumps(variants),
                     "peak_kind": "bf16 cuBLAS burst, " + peaks["source"],
                     "note": f"2*M*N*K summed over the step's Linears / CUDA-event time of the step ({n_calls} launches, one CUDA graph); traffic = DRAM "
                             "read+write bytes of the same launches from the committed ncu launch list (null until a list of this launch count "
                             "is committed); frac_per_shape_roofline = sum over launches of max(flops / burst peak, algorithmic bytes / "
                             "measured HBM) / step time"},
        "clocks": clocks,
    }
    if unfused is not None:
        line["unfused"] = unfused
    line.update(extra)
    if world == 1:
        line["cpu_baseline"] = {k: v for k, v in cpu_reference(20.0).items() if k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ per-layer / sweep tables
def run_tables(args):
    """--layers: per distinct layer shape of the workload; --sweep: BASELINE config 5 (W4A16 vs W8A8 vs bf16).
    Writes a JSON table to --out (default gpurun_out/); not part of the bench contract."""
    import torch

    q = importlib.import_module("quantization---diffusion-models_b200")
    shapes_mod, layers = layer_list()
    if args.model == "sdxl":      # BASELINE config 3: SDXL UNet, 1024^2, batch 4 + CFG
        layers = shapes_mod.sdxl_unet_linears(batch=4, cfg=True)
    elif args.model == "sd35":    # BASELINE config 4: SD3.5-L MMDiT, 1024^2, batch 1
        layers = shapes_mod.sd35_mmdit_linears(batch=1)
    if args.fused:                # the launches of the model with fused_utils.fuse_projections applied
        layers = [e[:5] for e in getattr(shapes_mod, {"sd15": "sd15_unet_linears_fused", "sdxl": "sdxl_unet_linears_fused",
                                                      "sd35": "sd35_mmdit_linears_fused"}[args.model])()]
    counts = {}
    for _, m_, n_, k_, c_ in layers:
        counts[(m_, n_, k_)] = counts.get((m_, n_, k_), 0) + c_
    dev = torch.device("cuda", 0)
    peaks = measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def graph_ms(body, reps=10):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            for _ in range(reps):
                body()
        g_.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g_.replay(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
        return best

    flush_ms = graph_ms(lambda: flush.zero_())

    def time_fn(fn):
        """GPU-side time of one launch: a CUDA graph of 10 x (256 MB L2 flush, launch) minus the flushes alone, so that
        neither host launch overhead (Python, ctypes, tensor-map encode) nor a warm L2 enters the number."""
        def body():
            flush.zero_()
            fn()
        if not args.sweep:
            return max(graph_ms(body) - flush_ms, 1e-4)
        # sweep: the large cases run seconds at the power cap and the case after them used to inherit the lowered clocks
        # (4096 x 6144 x 1536 right after 65536 x 6144 x 6144 read 0.37 for every kernel but cuBLAS): pause, best of two
        best = 1e9
        for _ in range(2):
            time.sleep(0.25)
            best = min(best, max(graph_ms(body) - flush_ms, 1e-4))
        return best

    if args.sweep:
        cases = []
        for m in (4096, 8192, 16384, 32768, 65536):
            for kn in (1536, 2048, 3072, 4096, 6144):
                cases.append((m, kn, kn))
        for m in (4096, 16384):
            cases += [(m, 6144, 1536), (m, 1536, 6144), (m, 8192, 2048), (m, 2048, 8192)]
    else:
        cases = sorted({(m, n, k) for _, m, n, k, _ in layers}, reverse=True)
    rows = []
    g = torch.Generator(device=dev).manual_seed(42)
    i8_peak = int8_dense_peak(torch, dev) if args.sweep else None     # the W8A8 denominator, this run
    if i8_peak:
        peaks = dict(peaks, int8_dense_tops=i8_peak, int8_how="torch._int_mm (cuBLASLt int8) 8192^3, best of 10, this run")
    for m, n, k in cases:
        grp = shapes_mod.group_for(k)
        x = torch.randn(m, k, generator=g, device=dev, dtype=torch.float16)
        w = torch.randn(n, k, generator=g, device=dev, dtype=torch.float16) * 0.02
        qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
        flops = 2.0 * m * n * k
        r = {"M": m, "N": n, "K": k, "group": grp, "calls_per_step": counts.get((m, n, k), 1)}
        bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp)     # the kernel-native copy WQLinear_GEMM keeps (dispatch as in the module)
        t = time_fn(lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, None, bts))
        r["w4a16_kernel"] = list(q.ops.gemm_last_variant())
        by = shapes_mod.gemm_bytes_w4a16(m, n, k, grp)
        roof = min(peaks["bf16_burst"], by and flops / by * peaks["hbm"] / 1e3)
        r["w4a16"] = {"ms": t, "tflops": flops / t / 1e9, "gbs": by / t / 1e6, "roof_tflops": roof, "frac": flops / t / 1e9 / roof}
        t = time_fn(lambda: q.ops.gemm_f16(x, dq))
        r["f16_tcgen05"] = {"ms": t, "tflops": flops / t / 1e9}
        t = time_fn(lambda: torch.nn.functional.linear(x, dq))
        r["cublas_f16"] = {"ms": t, "tflops": flops / t / 1e9}
        if args.sweep:    # config 5's third arm: the unquantised bf16 Linear (cuBLAS), the roofline denominator's own kernel
            xb, wb = x.bfloat16(), dq.bfloat16()
            t = time_fn(lambda: torch.nn.functional.linear(xb, wb))
            r["cublas_bf16"] = {"ms": t, "tflops": flops / t / 1e9, "frac": flops / t / 1e9 / peaks["bf16_burst"]}
            del xb, wb
        if k % 16 == 0:
            xq, sx = q.ops.actquant_token_i8(x)
            _, wq, sw, _ = q.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
            swf = sw.float()
            t = time_fn(lambda: q.ops.gemm_w8a8(xq, sx, wq, swf))
            r["w8a8_gemm"] = {"ms": t, "tflops": flops / t / 1e9}
            if i8_peak:   # roofline of kernel (d): min(measured int8 tensor peak, algorithmic bytes at the measured HBM rate)
                by8 = shapes_mod.gemm_bytes_w8a8(m, n, k)
                roof8 = min(i8_peak, flops / by8 * peaks["hbm"] / 1e3)
                r["w8a8_gemm"].update(roof_tops=roof8, frac=flops / t / 1e9 / roof8)
            t2 = time_fn(lambda: q.ops.actquant_token_i8(x))
            r["w8a8_actquant"] = {"ms": t2, "gbs": 3.0 * m * k / t2 / 1e6}
        rows.append(r)
        print(json.dumps(r), flush=True)
        del x, w, qw, qz, sc, dq, bts
    summary = None
    if not args.sweep:   # the whole Linear pass of one denoise step, every launch timed alone with a cold L2
        tot_f = sum(2.0 * r["M"] * r["N"] * r["K"] * r["calls_per_step"] for r in rows)
        summary = {"model": args.model, "tflop_per_step": tot_f / 1e12}
        for key, sub in (("w4a16", "w4a16"), ("w8a8", "w8a8_gemm"), ("f16_tcgen05", "f16_tcgen05"), ("cublas_f16", "cublas_f16")):
            if all(sub in r for r in rows):
                ms = sum(r[sub]["ms"] * r["calls_per_step"] for r in rows)
                summary[key] = {"ms_per_step": ms, "tflops": tot_f / ms / 1e9}
        if all("w8a8_actquant" in r for r in rows):
            ms = sum((r["w8a8_gemm"]["ms"] + r["w8a8_actquant"]["ms"]) * r["calls_per_step"] for r in rows)
            summary["w8a8_with_actquant"] = {"ms_per_step": ms, "tflops": tot_f / ms / 1e9}
        roof_ms = sum(2.0 * r["M"] * r["N"] * r["K"] / r["w4a16"]["roof_tflops"] / 1e9 * r["calls_per_step"] for r in rows)
        summary["w4a16_per_shape_roofline"] = {"ms_per_step": roof_ms, "frac": roof_ms / summary["w4a16"]["ms_per_step"]}
        print(json.dumps({"summary": summary}), flush=True)
    name = "gemm_sweep.json" if args.sweep else ("gemm_layers.json" if args.model == "sd15" else f"gemm_layers_{args.model}.json")
    out = args.out or os.path.join(ROOT, "gpurun_out", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"peaks": peaks, "summary": summary, "rows": rows}, f, indent=1)


# ------------------------------------------------------------------------------------------ model-level modes
def run_models(args):
    """The other BASELINE.json metrics, on the synthetic skeletons (one JSON line each; not the driver contract):
      --mode denoise : it/s of the 50-step style loop (CFG) with fp16 / W4A16 / W8A8 Linears, data parallel
      --mode calib   : AWQ calibration (scale + clip search, quantise, pack) seconds per model, block-sharded
      --mode rtn     : config 1, RTN W8 per-channel over every Linear / Conv weight of the SD1.5 UNet (GB/s)"""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M = importlib.import_module("quantization---diffusion-models_b200.models")
    q = importlib.import_module("quantization---diffusion-models_b200")
    cls = {"sd15": M.StableDiffusion1_x, "sdxl": M.StableDiffusionXL, "sd35": M.StableDiffusion3_5}[args.model]
    arch = {"layers": args.blocks} if (args.model == "sd35" and args.blocks) else {}
    batch = args.batch or {"sd15": 8, "sdxl": 4, "sd35": 1}[args.model]
    t_build = time.perf_counter()
    model = cls.from_skeleton(device=dev, **arch)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    def sync_max(x):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    out = {"mode": args.mode, "model": args.model, "n_gpus": world, "batch_per_gpu": batch, "data": "synthetic, random-init skeleton",
           "build_s": t_build}
    if args.mode == "denoise":
        cfgs = {"w4a16": {"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, "w8a8": {"w_bit": 8, "version": "w8a8"}}
        if args.quant != "fp16":
            model.quantize(quant_config=cfgs[args.quant], quantType="awq")
        prompts = [f"prompt {rank}-{i}" for i in range(batch)]
        lat = torch.randn(batch, model.pipeline.latent_channels, model.pipeline.latent_size, model.pipeline.latent_size,
                          generator=torch.Generator().manual_seed(42 + rank)).to(dev, model.pipeline.dtype)
        model.generate(prompts, lat=lat, num_inference_steps=2)
        sync_max(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q.ops.launch_count(reset=True)
        e0.record()
        res = model.generate(prompts, lat=lat, num_inference_steps=args.steps)
        e1.record()
        torch.cuda.synchronize()
        sec = sync_max(e0.elapsed_time(e1) * 1e-3)
        out.update({"metric": "denoise_it_per_s", "quant": args.quant, "steps": args.steps, "value": args.steps / sec,
                    "images_it_per_s": world * batch * args.steps / sec, "seconds": sec, "finite": bool(torch.isfinite(res).all()),
                    "libqdm_launches": q.ops.launch_count(), "scaling": "weak (prompt-batched data parallel, no collective in the loop)"})
    elif args.mode == "calib":
        model.calib_samples = model.default_calib_samples(args.calib_batches, batch)
        model.calib_steps = args.calib_steps
        _warm_pipeline(model)
        sync_max(0.0)
        t0 = time.perf_counter()
        model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq",
                       calibrate=True, shard=(rank, world) if world > 1 else None)
        torch.cuda.synchronize()
        sec = sync_max(time.perf_counter() - t0)
        chk, n_mod = 0, 0
        for name, mod in model.denoiser().named_modules():
            if type(mod).__name__ == "WQLinear_GEMM":
                n_mod += 1
                for t in (mod.qweight, mod.qzeros, mod.scales.view(torch.int16)):
                    chk = (chk * 1000003 + int(t.to(torch.int64).sum().item()) + t.numel()) % (1 << 61)
        out.update({"metric": "awq_calib_s_per_model", "value": sec, "blocks": len(model.get_search_blocks()), "groups_searched": len(model.quantizer.search_log),
                    "packed_modules": n_mod, "codes_checksum": chk, "phases": getattr(model.quantizer, "timings", None), "calib_batches": args.calib_batches, "calib_steps": args.calib_steps,
                    "note": "checksum covers qweight/qzeros/scales of every packed Linear; identical across world sizes = bit-exact codes"})
    else:  # rtn (config 1)
        fq = importlib.import_module("quantization---diffusion-models_b200.fake_quant")
        ws = [m.weight.data for m in model.denoiser().modules() if isinstance(m, (torch.nn.Linear, torch.nn.Conv2d))]
        numel = sum(w.numel() for w in ws)
        for w in ws[:4]:
            fq.quantize_weight_per_channel_absmax(w, 8)
        torch.cuda.synchronize()

        def rtn_pass():
            for w in ws:
                fq.quantize_weight_per_channel_absmax(w, 8)

        # 282 short launches: replayed as one CUDA graph so that the number is the kernels' and not Python's
        run_pass = rtn_pass
        if not args.no_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                rtn_pass()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                rtn_pass()
            run_pass = graph.replay
        run_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_pass()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        import oracle.qdm_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        sample = [w.cpu() for w in ws[:: max(1, len(ws) // 24)]]
        tc = time.perf_counter()
        for w in sample:
            O.rtn_rows(w, 8)
        tc = time.perf_counter() - tc
        cpu_numel = sum(w.numel() for w in sample)
        out.update({"metric": "rtn_w8_per_channel_GBps", "value": 4.0 * numel / sec / 1e9, "tensors": len(ws), "params": numel, "seconds": sec,
                    "bytes_model": "2 B read + 2 B written per weight (fake-quant output)",
                    "cpu_baseline": {"value": 4.0 * cpu_numel / tc / 1e9, "unit": "GB/s", "cores": torch.get_num_threads(), "kind": "port",
                                     "sample": f"{len(sample)} of {len(ws)} tensors, {cpu_numel / 1e6:.0f} M weights"}})
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_kernels(args):
    """--mode kernels: HBM roofline of the reduction (a) and quantise/pack (b) kernels on tensors larger than L2.
    Bytes are the algorithmic ones of SURVEY.md section 8(d): elem*numel read + the outputs actually emitted."""
    import torch
    q = importlib.import_module("quantization---diffusion-models_b200")
    dev = torch.device("cuda", 0)
    peaks = measured_peaks()
    g = torch.Generator(device=dev).manual_seed(0)
    rows, cols = 65536, 2560                      # 335 MB fp16 activations (one SD1.5 ff.net.0 output)
    x = torch.randn(rows, cols, generator=g, device=dev, dtype=torch.float16)
    w = torch.randn(9728 * 4, 2432, generator=g, device=dev, dtype=torch.float16) * 0.02   # 4 SD3.5 ff weights, 189 MB
    y = (x.float() + 0.01).half()
    s_vec = (torch.rand(2432, generator=g, device=dev) + 0.5).half()
    dq_out = torch.empty_like(w)

    def t_ms(fn, iters=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    nx, nw = x.numel(), w.numel()
    st_max = torch.zeros(cols, dtype=torch.float16, device=dev)
    st_acc1, st_acc2 = (torch.zeros(cols, dtype=torch.float64, device=dev) for _ in range(2))
    cases = [
        ("a colabsmax (hook, calib_data.py:117)", lambda: q.ops.colabsmax(x), 2 * nx),
        ("a colabssum (x_mean, quantizer.py:652)", lambda: q.ops.colabssum(x), 2 * nx),
        ("a colstats one pass: max, sum of maxima, |x| sum (calib_data.py:112-121)",
         lambda: q.ops.colstats(x, out_max=st_max, running=True, acc_maxsum=st_acc1, acc_abssum=st_acc2), 2 * nx),
        ("a awq_wsum (w_mean, quantizer.py:627)", lambda: q.ops.awq_wsum(w, 128), 2 * nw),
        ("a sqdiff_sum (loss, quantizer.py:777)", lambda: q.ops.sqdiff_sum(x, y), 4 * nx),
        ("a rowabsmax", lambda: q.ops.rowabsmax(x), 2 * nx),
        ("b quant_group zp g128 -> dq (quantizer.py:163)", lambda: q.ops.quant_group(w, 128, 4, True, want_scales=True, out=dq_out), 4 * nw + 4 * nw // 128),
        ("b quant_group W*s, /s fused (quantizer.py:727)", lambda: q.ops.quant_group(w, 128, 4, True, pre_mul=s_vec, post_div=s_vec, want_scales=False, out=dq_out), 4 * nw),
        ("b quant_group sym noclamp (fake_quant.py:21)", lambda: q.ops.quant_group(w, 128, 4, False, no_clamp=True, want_scales=False, out=dq_out), 4 * nw),
        ("b quant_pack_awq (quantizer.py:540-569)", lambda: q.ops.quant_pack_awq(w, 128), 2 * nw + nw // 2 + 5 * nw // 256),
        ("b quant_rowwise 8-bit (fake_quant.py:86)", lambda: q.ops.quant_rowwise(x, 8), 4 * nx),
        ("b actquant_token_i8 (fake_quant.py:109)", lambda: q.ops.actquant_token_i8(x), 3 * nx + 4 * rows),
        ("b dequant_awq (packing_utils.py:87)", None, 0),
        ("act geglu [65536, 2560] -> [65536, 1280] (ff.net.0.proj -> ff.net.2)", lambda: q.ops.geglu(x), 3 * nx),
        ("ref torch h * F.gelu(gate) on the same tensor", lambda: x[:, :cols // 2] * torch.nn.functional.gelu(x[:, cols // 2:]), 3 * nx),
        ("ref torch copy_ (same bytes model: 2 B in + 2 B out)", lambda: dq_out.copy_(w), 4 * nw),
    ]
    qw, qz, sc, _ = q.ops.quant_pack_awq(w, 128)
    cases[[c[0] for c in cases].index("b dequant_awq (packing_utils.py:87)")] = ("b dequant_awq (packing_utils.py:87)", lambda: q.ops.dequant_awq(qw, qz, sc, 128), nw // 2 + 2 * nw + 5 * nw // 256)
    rows_out = []
    for name, fn, nbytes in cases:
        ms = t_ms(fn)
        r = {"kernel": name, "ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peaks["hbm"]}
        rows_out.append(r)
        print(json.dumps(r), flush=True)
    out = args.out or os.path.join(ROOT, "gpurun_out", "kernels_ab.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"peaks": peaks, "rows": rows_out}, f, indent=1)


def run_conv(args):
    """--mode conv: the 3x3 convolutions of the SD1.5 UNet (batch 8 + CFG) as implicit GEMMs on kernels (c) / f16 against
    cuDNN's fp16 conv2d on the same tensors.  TFLOP/s = 2 * B*H*W * N * 9C / t.  `ours_*`: direct form (4-D TMA, no padded
    copy) on channels-last input; `_from_nchw`: the same with the NCHW -> NHWC transpose pass inside the timed call;
    `_padded_grid`: the fallback form with its zero-pad pass.  L2 flushed before every call."""
    import torch
    q = importlib.import_module("quantization---diffusion-models_b200")
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    B = 16
    shapes = [(320, 320, 64), (640, 320, 64), (960, 320, 64), (640, 640, 32), (960, 640, 32), (1280, 640, 32), (1920, 640, 32),
              (1280, 1280, 16), (1920, 1280, 16), (2560, 1280, 16), (1280, 1280, 8), (2560, 1280, 8)]

    def t_ms(fn, iters=5):
        for _ in range(2):
            fn()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    rows = []
    for C, N, H in shapes:
        x = torch.randn(B, C, H, H, generator=g, device=dev, dtype=torch.float16)
        x_cl = x.contiguous(memory_format=torch.channels_last)
        w = torch.randn(N, C, 3, 3, generator=g, device=dev, dtype=torch.float16) * 0.03
        w_cl = w.contiguous(memory_format=torch.channels_last)
        b = torch.randn(N, generator=g, device=dev, dtype=torch.float16)
        taps = q.ops.conv3x3_weight_taps(w)
        grp = 128 if (9 * C) % 128 == 0 else 64
        qw, qz, sc, _ = q.ops.quant_pack_awq(taps, grp)
        flops = 2.0 * B * H * H * N * 9 * C
        r = {"C": C, "N": N, "H": H, "gflop": flops / 1e9,
             "cudnn_nchw_ms": t_ms(lambda: torch.nn.functional.conv2d(x, w, b, 1, 1)),
             "cudnn_nhwc_ms": t_ms(lambda: torch.nn.functional.conv2d(x_cl, w_cl, b, 1, 1)),
             "ours_f16_ms": t_ms(lambda: q.ops.conv3x3_f16(x_cl, taps, b)),
             "ours_f16_padded_grid_ms": t_ms(lambda: q.ops.conv3x3_f16(x_cl, taps, b, padded=True)),
             "ours_f16_from_nchw_ms": t_ms(lambda: q.ops.conv3x3_f16(x, taps, b)),
             "ours_w4a16_ms": t_ms(lambda: q.ops.conv3x3_w4a16(x_cl, qw, qz, sc, grp, b))}
        for k in list(r):
            if k.endswith("_ms"):
                r[k[:-3] + "_tflops"] = flops / r[k] / 1e9
        rows.append(r)
        print(json.dumps(r), flush=True)
    # the three stride-2 down-samplers (Downsample2D: 3x3, stride 2, padding 1) of the same UNet: qdm_conv3x3s2_nhwc_*
    for C, H in ((320, 64), (640, 32), (1280, 16)):
        x = torch.randn(B, C, H, H, generator=g, device=dev, dtype=torch.float16)
        x_cl = x.contiguous(memory_format=torch.channels_last)
        w = torch.randn(C, C, 3, 3, generator=g, device=dev, dtype=torch.float16) * 0.03
        w_cl = w.contiguous(memory_format=torch.channels_last)
        b = torch.randn(C, generator=g, device=dev, dtype=torch.float16)
        taps = q.ops.conv3x3_weight_taps(w)
        grp = 128 if (9 * C) % 128 == 0 else 64
        qw, qz, sc, _ = q.ops.quant_pack_awq(taps, grp)
        flops = 2.0 * B * (H // 2) * (H // 2) * C * 9 * C
        r = {"C": C, "N": C, "H": H, "stride": 2, "gflop": flops / 1e9,
             "cudnn_nchw_ms": t_ms(lambda: torch.nn.functional.conv2d(x, w, b, 2, 1)),
             "cudnn_nhwc_ms": t_ms(lambda: torch.nn.functional.conv2d(x_cl, w_cl, b, 2, 1)),
             "ours_f16_ms": t_ms(lambda: q.ops.conv3x3_f16(x_cl, taps, b, stride=2)),
             "ours_f16_from_nchw_ms": t_ms(lambda: q.ops.conv3x3_f16(x, taps, b, stride=2)),
             "ours_w4a16_ms": t_ms(lambda: q.ops.conv3x3_w4a16(x_cl, qw, qz, sc, grp, b, stride=2))}
        for k in list(r):
            if k.endswith("_ms"):
                r[k[:-3] + "_tflops"] = flops / r[k] / 1e9
        rows.append(r)
        print(json.dumps(r), flush=True)
    out = args.out or os.path.join(ROOT, "gpurun_out", "conv3x3.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"batch": B, "rows": rows}, f, indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--layers", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-fuse", action="store_true", help="every Linear as its own launch (no same-input fusion)")
    ap.add_argument("--no-chain", action="store_true", help="e2e: ff.net.2 inputs from the host too (no on-device GEGLU)")
    ap.add_argument("--fused", action="store_true", help="--layers: the fused launch inventory of the model")
    ap.add_argument("--no-extras", action="store_true", help="skip the w8a8 / denoise / calib sub-records of the line")
    ap.add_argument("--extras", default="w8a8,config5,denoise,calib", help="which sub-records to run (comma-separated)")
    ap.add_argument("--mode", default="linears", choices=["linears", "denoise", "calib", "rtn", "kernels", "conv"])
    ap.add_argument("--model", default="sd15", choices=["sd15", "sdxl", "sd35"])
    ap.add_argument("--quant", default="w4a16", choices=["fp16", "w4a16", "w8a8"])
    ap.add_argument("--blocks", type=int, default=0, help="sd35: number of joint blocks (default 38)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--calib-batches", type=int, default=1)
    ap.add_argument("--calib-steps", type=int, default=2)
    args = ap.parse_args()
    if args.mode == "kernels":
        return run_kernels(args)
    if args.mode == "conv":
        return run_conv(args)
    if args.mode != "linears":
        return run_models(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.sweep or args.layers:
        return run_tables(args)
    run_ours(args)


if __name__ == "__main__":
    main()
