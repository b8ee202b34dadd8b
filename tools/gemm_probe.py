"""GPU bring-up probe for the tcgen05 GEMMs: each case runs in its own subprocess (a trapped kernel kills
only that case) and prints the error plus a coarse map of where the output is wrong.

    python tools/gemm_probe.py            # all cases
    python tools/gemm_probe.py one <kind> <dt> <M> <N> <K> [group]
"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    ("f16", "f16", 128, 128, 64), ("f16", "f16", 128, 128, 256), ("f16", "bf16", 256, 256, 128),
    ("f16", "f16", 300, 320, 320), ("f16", "f16", 1024, 2432, 2432),
    ("f16_kn", "f16", 128, 128, 64), ("f16_kn", "f16", 128, 256, 128), ("f16_kn", "bf16", 300, 320, 320),
    ("w4", "f16", 128, 128, 128, 128), ("w4", "f16", 128, 256, 256, 64), ("w4", "bf16", 300, 320, 320, 64),
    ("w4", "f16", 1024, 2432, 2432, 128), ("w4", "f16", 4096, 2560, 320, 64),
    ("w8", "f16", 128, 128, 128), ("w8", "f16", 300, 320, 320), ("w8", "bf16", 1024, 1280, 1280),
    ("f16", "f16", 4096, 1280, 1280), ("w4", "f16", 8192, 1280, 1280, 128), ("w4", "bf16", 4096, 640, 2560, 128),
    ("w8", "f16", 8192, 5120, 640), ("w4", "f16", 257, 256, 128, 128), ("w4", "f16", 65536, 320, 320, 64),
    ("w4", "f16", 16, 1280, 1280, 128), ("w4", "bf16", 2, 14592, 2432, 128), ("w4", "f16", 1, 320, 1280, 128),
    ("w4", "f16", 9, 1280, 320, 64), ("w4", "f16", 16, 72, 192, 64),
]


def run_one(kind, dt, M, N, K, group=128):
    import torch
    q = importlib.import_module("quantization---diffusion-models_b200")
    dtype = {"f16": torch.float16, "bf16": torch.bfloat16}[dt]
    dev = "cuda:0"
    q.ops.set_gemm_mode(int(os.environ.get("QDM_GEMM_MODE", "0")))
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dtype).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(dtype).to(dev)
    b = torch.randn(N, generator=g).to(dtype).to(dev)
    if kind == "f16":
        y = q.ops.gemm_f16(x, w, b)
        ref = x.float() @ w.float().t() + b.float()
    elif kind == "f16_kn":
        y = q.ops.gemm_f16_kn(x, w.t().contiguous(), b)
        ref = x.float() @ w.float().t() + b.float()
    elif kind == "w4":
        qw, qz, sc, dq = q.ops.quant_pack_awq(w, group, want_dq=True)
        y = q.ops.gemm_w4a16(x, qw, qz, sc, group, b)
        ref = x.float() @ dq.float().t() + b.float()
    else:
        xq, sx = q.ops.actquant_token_i8(x)
        _, wq, sw, _ = q.ops.quant_rowwise(w, 8, want_dq=False, want_codes=True, want_scales=True)
        y = q.ops.gemm_w8a8(xq, sx, wq, sw.float(), b, out_dtype=dtype)
        ref = (xq.float() @ wq.float().t()) * sx[:, None] * sw.float()[None, :] + b.float()
    torch.cuda.synchronize()
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
    print(f"{kind} {dt} M={M} N={N} K={K} g={group}: max rel err {err:.3e}", flush=True)
    if not (err < 1e-2):
        bad = ((y.float() - ref).abs() > 0.02 * ref.abs().max())
        print("  bad fraction", bad.float().mean().item())
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:32].tolist()
        print("  first bad rows", rows)
        print("  first bad cols", cols)
        print("  y[0,:8]  ", y[0, :8].float().tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        # per 8-column block error map of the first 128x128 corner
        blk = bad[:128, :128].float().reshape(min(M, 128) // 8 if M >= 8 else 1, -1, min(N, 128) // 8, 8).mean(dim=(1, 3)) if M >= 128 else None
        if blk is not None:
            for r in range(blk.shape[0]):
                print("  " + "".join("#" if v > 0.5 else ("+" if v > 0 else ".") for v in blk[r].tolist()))
        sys.exit(3)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        a = sys.argv[2:]
        run_one(a[0], a[1], int(a[2]), int(a[3]), int(a[4]), int(a[5]) if len(a) > 5 else 128)
        return
    fails = 0
    for c in CASES:
        cmd = [sys.executable, os.path.abspath(__file__), "one"] + [str(v) for v in c]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
            out = (r.stdout + r.stderr[-1500:]) if r.returncode else r.stdout
            print(out.rstrip(), flush=True)
            if r.returncode:
                print(f"  -> exit {r.returncode}", flush=True)
                fails += 1
        except subprocess.TimeoutExpired:
            print(f"{c}: TIMEOUT", flush=True)
            fails += 1
    print("FAILS", fails)


if __name__ == "__main__":
    main()
