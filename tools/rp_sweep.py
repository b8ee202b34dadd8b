#!/usr/bin/env python
"""A/B timing of the W4A16 kernels on given shapes (GPU-side time: CUDA graph of 10 x (256 MB L2 flush, launch) minus the
flushes alone, best of 3).  For every shape: the AWQ-tensor kernels (mode 64), the repacked-weight kernel with one / two
sub-tiles at every sub-tile width, and what the dispatcher picks on its own.  Used to calibrate choose_rp (qdm_gemm.cu).

    python tools/rp_sweep.py [--model sd15|sdxl|sd35|sweep|mid] [--out gpurun_out/rp_sweep.json] [--widths 64,128,160,256]
"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mid")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "rp_sweep.json"))
    ap.add_argument("--widths", default="")
    ap.add_argument("--dtype", default="f16")
    args = ap.parse_args()
    q = importlib.import_module("quantization---diffusion-models_b200")
    S = importlib.import_module("quantization---diffusion-models_b200.shapes")
    dev = torch.device("cuda", 0)
    dt = torch.float16 if args.dtype == "f16" else torch.bfloat16
    if args.model == "sd15":
        cases = sorted({(m, n, k) for _, m, n, k, _ in S.sd15_unet_linears(8, True) if m > 128}, reverse=True)
    elif args.model == "sdxl":
        cases = sorted({(m, n, k) for _, m, n, k, _ in S.sdxl_unet_linears(4, True) if m > 128}, reverse=True)
    elif args.model == "sd35":
        cases = sorted({(m, n, k) for _, m, n, k, _ in S.sd35_mmdit_linears(1) if m > 128}, reverse=True)
    elif args.model == "sweep":
        cases = [(4096, 4096, 4096), (8192, 4096, 4096), (16384, 2048, 2048), (65536, 1536, 1536), (4096, 8192, 2048), (4096, 2048, 8192)]
    else:
        cases = [(4096, 1280, 1280), (16384, 640, 640), (65536, 320, 320), (4096, 10240, 1280), (16384, 5120, 640), (65536, 2560, 320),
                 (4096, 1280, 5120), (1232, 1280, 768), (1024, 1280, 1280), (8192, 1280, 1280), (4096, 2432, 2432), (333, 2432, 2432)]
    widths = [int(w) for w in args.widths.split(",")] if args.widths else [64, 96, 128, 160, 192, 224, 256]
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

    rows = []
    g = torch.Generator(device=dev).manual_seed(42)
    for m, n, k in cases:
        grp = S.group_for(k)
        x = torch.randn(m, k, generator=g, device=dev, dtype=dt)
        w = torch.randn(n, k, generator=g, device=dev, dtype=dt) * 0.02
        qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
        blob = q.ops.w4a16_repack(qw, qz, sc, grp)
        flops = 2.0 * m * n * k
        r = {"M": m, "N": n, "K": k, "group": grp}
        try:
            q.ops.set_gemm_mode(64)
            t = time_fn(lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, blob))
            r["awq"] = {"us": t * 1e3, "tflops": flops / t / 1e9, "variant": q.ops.gemm_last_variant()}
            q.ops.set_gemm_mode(0)
            t = time_fn(lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, blob))
            r["auto"] = {"us": t * 1e3, "tflops": flops / t / 1e9, "variant": q.ops.gemm_last_variant()}
            for mode, name in ((16, "rp1"), (32, "rp2")):
                for wd in widths:
                    if name == "rp2" and wd >= n:
                        continue
                    if name == "rp1" and wd - 32 >= n:
                        continue
                    q.ops.set_gemm_mode(mode | (wd << 8))
                    t = time_fn(lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, blob))
                    r[f"{name}_{wd}"] = round(t * 1e3, 2)
            q.ops.set_gemm_mode(0)
            t = time_fn(lambda: torch.nn.functional.linear(x, dq))
            r["cublas_f16"] = {"us": t * 1e3, "tflops": flops / t / 1e9}
        finally:
            q.ops.set_gemm_mode(0)
        rows.append(r)
        print(json.dumps(r), flush=True)
        del x, w, qw, qz, sc, dq, blob
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
