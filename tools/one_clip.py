"""Profile the AWQ clip search kernels on ONE weight: python tools/one_clip.py <co> <ci> [n_tok] [iters]
(qdm_awq_clip_search: per-group Gram matrix of the calibration rows + the 10-level search; used under ncu)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
q = importlib.import_module("quantization---diffusion-models_b200")


def main():
    co, ci = int(sys.argv[1]), int(sys.argv[2])
    n_tok = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    w = torch.randn(co, ci, generator=g, device=dev, dtype=torch.float16) * 0.02
    x = torch.randn(n_tok, ci, generator=g, device=dev, dtype=torch.float16)
    for _ in range(2):
        q.ops.awq_clip_search(w, x, 128)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        q.ops.awq_clip_search(w, x, 128)
    e1.record()
    torch.cuda.synchronize()
    print(f"clip search {co} x {ci}, {n_tok} rows: {e0.elapsed_time(e1) / iters * 1e3:.1f} us per call")


if __name__ == "__main__":
    main()
