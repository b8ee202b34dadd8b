"""Launch the HBM-bound kernels (a)/(b) + qdm_geglu once as warm-up and once more (for `ncu -k regex:... --launch-skip 5 -c 5`):
awq_wsum, quant_pack_awq, dequant_awq, fused-search quant_group, geglu -- on the tensors of `bench.py --mode kernels`."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("quantization---diffusion-models_b200")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(65536, 2560, generator=g, device=dev, dtype=torch.float16)
w = torch.randn(9728 * 4, 2432, generator=g, device=dev, dtype=torch.float16) * 0.02
s_vec = (torch.rand(2432, generator=g, device=dev) + 0.5).half()
dq_out = torch.empty_like(w)
qw, qz, sc, _ = q.ops.quant_pack_awq(w, 128)
torch.cuda.synchronize()
for _ in range(2):
    q.ops.awq_wsum(w, 128)
    q.ops.quant_pack_awq(w, 128)
    q.ops.dequant_awq(qw, qz, sc, 128)
    q.ops.quant_group(w, 128, 4, True, pre_mul=s_vec, post_div=s_vec, want_scales=False, out=dq_out)
    q.ops.geglu(x)
    torch.cuda.synchronize()
print("ok")
