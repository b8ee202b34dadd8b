"""Time / profile ONE GEMM shape: python tools/one_gemm.py <kind> <M> <N> <K> [iters]   (kind: f16|f16_kn|w4|w8|cublas)
Used under ncu for the per-kernel captures in profiles/ (keeps the report small)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
q = importlib.import_module("quantization---diffusion-models_b200")
sh = importlib.import_module("quantization---diffusion-models_b200.shapes")


def main():
    kind, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(M, K, generator=g, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, generator=g, device=dev, dtype=torch.float16) * 0.02
    grp = sh.group_for(K)
    if kind == "w4":
        qw, qz, sc, _ = q.ops.quant_pack_awq(w, grp)
        fn = lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp)
    elif kind == "w4rp":   # repacked-weight kernel; QDM_GEMM_MODE = 16 | 32 (+ width << 8) pins the form
        qw, qz, sc, _ = q.ops.quant_pack_awq(w, grp)
        blob = q.ops.w4a16_repack(qw, qz, sc, grp)
        fn = lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, blob)
    elif kind == "w4ts":   # TMEM-A kernel; QDM_GEMM_MODE = 128 (+ tokens per tile << 8)
        qw, qz, sc, _ = q.ops.quant_pack_awq(w, grp)
        bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp)
        fn = lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, None, bts)
    elif kind == "f16":
        fn = lambda: q.ops.gemm_f16(x, w)
    elif kind == "f16_kn":
        wk = w.t().contiguous()
        fn = lambda: q.ops.gemm_f16_kn(x, wk)
    elif kind == "w8":
        xq, sx = q.ops.actquant_token_i8(x)
        _, wq, sw, _ = q.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
        swf = sw.float()
        fn = lambda: q.ops.gemm_w8a8(xq, sx, wq, swf)
    else:
        fn = lambda: torch.nn.functional.linear(x, w)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    if os.environ.get("QDM_ONE_GRAPH"):
        # warm GPU-side time without host launch overhead: CUDA graph of 10 x (L2 flush, gemm) minus 10 x flush
        def graph_ms(body):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                for _ in range(10):
                    body()
            g_.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g_.replay(); e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best / 10
        t_both = graph_ms(lambda: (flush.zero_(), fn()))
        t_flush = graph_ms(lambda: flush.zero_())
        t = t_both - t_flush
        print(f"{kind} M={M} N={N} K={K}: {t * 1e3:.1f} us GPU-side (graph, L2 flushed), {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s")
        return
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    print(f"{kind} M={M} N={N} K={K}: {t * 1e3:.1f} us, {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s (median of {iters}, L2 flushed)")


if __name__ == "__main__":
    main()
