#!/usr/bin/env python
"""Bring-up probe of the TMEM-A W4A16 kernel (qdm_gemm_w4ts.cu): correctness against the fp32 reference and GPU-side time
against the other W4 kernels and cuBLAS, per shape and token-tile width.   python tools/ts_probe.py [check|time|quick|models|fused|sweep] [dtype]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
q = importlib.import_module("quantization---diffusion-models_b200")
S = importlib.import_module("quantization---diffusion-models_b200.shapes")
dev = "cuda:0"


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "check"
    dt = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float16
    g = torch.Generator(device=dev).manual_seed(1)
    if what == "check":
        cases = [(256, 256, 128), (256, 256, 256), (512, 512, 512), (300, 320, 320), (777, 520, 448), (4096, 1280, 1280), (1232, 1280, 768),
                 (333, 2432, 2432), (4096, 64, 2432), (8192, 640, 2560), (129, 72, 256), (200, 8, 128)]
        tiles = [0, 384, 320, 256, 192, 128, 64]
    elif what == "quick":
        cases = [(4096, 10240, 1280), (4096, 1280, 1280), (16384, 5120, 640), (4096, 9728, 2432), (1232, 1280, 768), (16384, 640, 640)]
        tiles = [0, 320, 192]
    elif what == "models":   # every distinct Linear shape (M > 32) of the three denoisers
        seen = []
        for layers in (S.sd15_unet_linears(batch=8, cfg=True), S.sdxl_unet_linears(batch=4, cfg=True), S.sd35_mmdit_linears(batch=1)):
            for _, m_, n_, k_, _ in layers:
                if m_ > 32 and (m_, n_, k_) not in seen:
                    seen.append((m_, n_, k_))
        cases = seen
        tiles = [0] if os.environ.get("TS_PROBE_AUTO") else [0, 384, 320, 256, 192, 160, 128, 96, 64]
    elif what == "fused":   # the launch shapes fused_utils.fuse_projections adds (M > 32)
        base, seen = set(), []
        for layers in (S.sd15_unet_linears(), S.sdxl_unet_linears(), S.sd35_mmdit_linears()):
            base |= {(m_, n_, k_) for _, m_, n_, k_, _ in layers}
        for layers in (S.sd15_unet_linears_fused(), S.sdxl_unet_linears_fused(), S.sd35_mmdit_linears_fused()):
            for e in layers:
                c = (e[1], e[2], e[3])
                if e[1] > 32 and c not in base and c not in seen:
                    seen.append(c)
        cases = seen
        tiles = [0] if os.environ.get("TS_PROBE_AUTO") else [0, 384, 320, 256, 192, 160, 128]
    elif what == "sweep":    # BASELINE config 5 corners: long-running problems at the power cap
        cases = [(32768, 4096, 4096), (65536, 3072, 3072), (16384, 6144, 6144), (65536, 1536, 1536), (16384, 2048, 8192), (4096, 2048, 2048)]
        tiles = [0, 384, 320, 256, 192, 160]
    elif what == "sweep_small":    # the config-5 shapes below 0.70 of the bf16 peak
        cases = [(4096, 1536, 1536), (4096, 2048, 2048), (8192, 1536, 1536), (4096, 3072, 3072), (8192, 2048, 2048)]
        tiles = [0, 384, 320, 256, 192, 160, 128]
    else:
        cases = [(4096, 10240, 1280), (4096, 1280, 1280), (16384, 5120, 640), (16384, 640, 640), (4096, 1280, 5120), (8192, 1280, 1280),
                 (4096, 2432, 2432), (4096, 9728, 2432), (1232, 1280, 768), (333, 2432, 2432), (65536, 2560, 320), (65536, 320, 320)]
        tiles = [0, 192, 160, 128, 96]
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

    flush_ms = graph_ms(lambda: flush.zero_()) if what != "check" else 0.0
    for m, n, k in cases:
        grp = S.group_for(k)
        x = torch.randn(m, k, generator=g, device=dev, dtype=dt)
        w = (torch.randn(n, k, generator=g, device=dev) * 0.05).to(dt)
        b = torch.randn(n, generator=g, device=dev).to(dt)
        if n % 64 == 0:
            qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
        else:
            import oracle.qdm_oracle as O
            oq, oz, os_, dq = O.awq_from_linear(w.cpu(), grp, 4)
            qw, qz, sc, dq = torch.from_numpy(oq).to(dev), torch.from_numpy(oz).to(dev), os_.to(dev), dq.to(dev)
        bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp)
        ref = x.float() @ dq.float().t() + b.float()
        row = f"{m:6d} {n:6d} {k:6d}"
        try:
            if what != "check":
                q.ops.set_gemm_mode(64)
                t = graph_ms(lambda: (flush.zero_(), q.ops.gemm_w4a16(x, qw, qz, sc, grp, b))) - flush_ms
                row += f" | awq {t * 1e3:7.1f} us {q.ops.gemm_last_variant()}"
                t = graph_ms(lambda: (flush.zero_(), torch.nn.functional.linear(x, dq, b))) - flush_ms
                row += f" | cublas {t * 1e3:7.1f}"
            for tl in tiles:
                q.ops.set_gemm_mode(128 | (tl << 8))
                y = q.ops.gemm_w4a16(x, qw, qz, sc, grp, b, None, bts)
                v = q.ops.gemm_last_variant()
                torch.cuda.synchronize()
                err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
                if what != "check":
                    t = graph_ms(lambda: (flush.zero_(), q.ops.gemm_w4a16(x, qw, qz, sc, grp, b, None, bts))) - flush_ms
                    row += f" | ts{v[1]} {t * 1e3:6.1f} ({2.0 * m * n * k / t / 1e9:5.0f} TF) e{err:.0e}"
                else:
                    row += f" | {v[0]}{v[1]} err {err:.2e}"
        finally:
            q.ops.set_gemm_mode(0)
        print(row, flush=True)


if __name__ == "__main__":
    main()
