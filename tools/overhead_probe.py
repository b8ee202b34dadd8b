"""Fixed cost of a launch per kernel family: GPU-side time per launch of tiny / small problems, back to back in one CUDA
graph (programmatic dependent launch chain, warm L2) and with an L2 flush between launches.
    python tools/overhead_probe.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
q = importlib.import_module("quantization---diffusion-models_b200")
S = importlib.import_module("quantization---diffusion-models_b200.shapes")
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def graph_ms(body, reps):
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
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g_.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def main():
    g = torch.Generator(device=dev).manual_seed(3)
    flush_ms = graph_ms(lambda: flush.zero_(), 10)
    print(f"flush alone {flush_ms * 1e3:.1f} us")
    cases = [(64, 256, 128), (192, 256, 128), (192, 256, 1280), (1232, 1280, 768), (4096, 1280, 1280), (16, 1280, 1280), (16, 1280, 320), (16, 320, 1280),
             (512, 320, 320), (16384, 320, 320), (65536, 320, 320)]
    for m, n, k in cases:
        grp = S.group_for(k)
        x = torch.randn(m, k, generator=g, device=dev, dtype=torch.float16)
        w = (torch.randn(n, k, generator=g, device=dev) * 0.05).half()
        qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
        bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp)
        row = f"{m:6d} {n:5d} {k:5d}"
        for name, fn in (("w4a16", lambda: q.ops.gemm_w4a16(x, qw, qz, sc, grp, None, None, bts)), ("cublas", lambda: torch.nn.functional.linear(x, dq))):
            fn()
            v = q.ops.gemm_last_variant() if name == "w4a16" else ""
            chain = graph_ms(fn, 20)
            cold = graph_ms(lambda: (flush.zero_(), fn()), 10) - flush_ms
            row += f" | {name} {str(v):14s} chain {chain * 1e3:6.2f} us  cold {cold * 1e3:6.2f} us"
        print(row, flush=True)


if __name__ == "__main__":
    main()
