"""One small launch of every kernel family, for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_probe.py`
(results recorded in profiles/sanitizer_r02.txt).  Each result is also checked against the fp32 reference."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
q = importlib.import_module("quantization---diffusion-models_b200")
S = importlib.import_module("quantization---diffusion-models_b200.shapes")
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(5)


def w4(m, n, k, mode=0, ts=True, dt=torch.float16):
    grp = S.group_for(k)
    x = torch.randn(m, k, generator=g, device=dev, dtype=dt)
    w = (torch.randn(n, k, generator=g, device=dev) * 0.05).to(dt)
    b = torch.randn(n, generator=g, device=dev).to(dt)
    qw, qz, sc, dq = q.ops.quant_pack_awq(w, grp, want_dq=True)
    bts = q.ops.w4a16_repack_ts(qw, qz, sc, grp) if ts else None
    q.ops.set_gemm_mode(mode)
    y = q.ops.gemm_w4a16(x, qw, qz, sc, grp, b, None, bts)
    q.ops.set_gemm_mode(0)
    ref = x.float() @ dq.float().t() + b.float()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    print(f"w4a16 {m}x{n}x{k} mode {mode} -> {q.ops.gemm_last_variant()} err {err:.1e}", flush=True)
    assert err < 1e-2


def main():
    w4(300, 512, 256)                      # TS, two token tiles, K = 256
    w4(1232, 1280, 768, dt=torch.bfloat16) # TS bf16
    w4(4096, 320, 320)                     # TS with a partly empty channel block
    w4(512, 640, 640, mode=2, ts=False)    # CTA-pair kernel, shared-memory B
    w4(65536, 320, 320, ts=False)          # B-stationary kernel
    w4(4096, 1280, 2560, mode=8, ts=False) # stream-K
    w4(100, 256, 128, mode=1, ts=False)    # single-CTA kernel
    w4(16, 1280, 320)                      # small-M mma.sync kernel
    w4(1, 4864, 2432)                      # skinny cluster split-K kernel
    # W8A8 and f16 tcgen05 GEMMs
    x = torch.randn(777, 640, generator=g, device=dev, dtype=torch.float16)
    w = (torch.randn(320, 640, generator=g, device=dev) * 0.05).half()
    xq, sx = q.ops.actquant_token_i8(x)
    _, wq, sw, _ = q.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
    y = q.ops.gemm_w8a8(xq, sx, wq, sw.float())
    ref = (xq.float() @ wq.float().t()) * sx[:, None] * sw.float()[None, :]
    print("w8a8 err", ((y.float() - ref).abs().max() / ref.abs().max()).item(), flush=True)
    y = q.ops.gemm_f16(x, w)
    print("f16 err", ((y.float() - x.float() @ w.float().t()).abs().max() / y.float().abs().max()).item(), flush=True)
    # quantise / reduce / clip kernels
    q.ops.quant_group(w, 128, 4, True)
    q.ops.colabsmax(x); q.ops.colabssum(x); q.ops.awq_wsum(w, 128)
    q.ops.awq_clip_search(w, x[:256], 128)
    torch.cuda.synchronize()
    print("sanitize probe done", flush=True)


if __name__ == "__main__":
    main()
