"""One implicit-GEMM 3x3 convolution for ncu:  python tools/one_conv.py {f16|w4} C N H [B] [padded]
    ncu --set full --clock-control none --import-source on -k regex:qdm_gemm2 -c 1 -o gpurun_out/prof_conv python tools/one_conv.py w4 640 640 32"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("quantization---diffusion-models_b200")
kind, C, N, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 16
padded = True if (len(sys.argv) > 6 and sys.argv[6] == "padded") else None
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, C, H, H, generator=g, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
w = torch.randn(N, C, 3, 3, generator=g, device=dev, dtype=torch.float16) * 0.03
taps = q.ops.conv3x3_weight_taps(w)
grp = 128 if (9 * C) % 128 == 0 else 64
qw, qz, sc, _ = q.ops.quant_pack_awq(taps, grp)
for _ in range(3):
    y = q.ops.conv3x3_f16(x, taps, None, padded=padded) if kind == "f16" else q.ops.conv3x3_w4a16(x, qw, qz, sc, grp, None, padded=padded)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
