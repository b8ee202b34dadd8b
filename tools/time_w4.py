"""GPU-side time of qdm_gemm_w4a16 on a few (M, N, K) shapes: a CUDA graph of 10 x (256 MB L2 flush, launch) minus the
flushes alone (the method of bench.py --layers).   python tools/time_w4.py 1x14592x2432 16x1280x1280 ..."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("quantization---diffusion-models_b200")
dev = torch.device("cuda", 0)
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
g = torch.Generator(device=dev).manual_seed(42)
for spec in sys.argv[1:]:
    m, n, k = (int(v) for v in spec.split("x"))
    grp = 128 if k % 128 == 0 else 64
    x = torch.randn(m, k, generator=g, device=dev, dtype=torch.float16)
    w = torch.randn(n, k, generator=g, device=dev, dtype=torch.float16) * 0.02
    qw, qz, sc, _ = q.ops.quant_pack_awq(w, grp)

    def body():
        flush.zero_()
        q.ops.gemm_w4a16(x, qw, qz, sc, grp)

    t = max(graph_ms(body) - flush_ms, 1e-4)
    print(json.dumps({"M": m, "N": n, "K": k, "us": t * 1e3, "packed_GBps": (0.5 * n * k + 2.5 * (k // grp) * n) / t / 1e6}), flush=True)
