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
