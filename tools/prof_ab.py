"""One launch of every (a)/(b) kernel on an HBM-sized tensor, for `ncu --set full` (see profiles/).
    ncu --set full --clock-control none -k regex:'quant_|awq_wsum_stage1|sqdiff_stage1|col_reduce_stage1|col_stats_stage1|dequant_awq' ..."""
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
y = (x.float() + 0.01).half()
s_vec = (torch.rand(2432, generator=g, device=dev) + 0.5).half()
out = torch.empty_like(w)
st_max = torch.zeros(2560, dtype=torch.float16, device=dev)
st_acc, st_acc2 = (torch.zeros(2560, dtype=torch.float64, device=dev) for _ in range(2))
for _ in range(2):
    q.ops.quant_group(w, 128, 4, True, want_scales=True, out=out)
    q.ops.quant_group(w, 128, 4, True, pre_mul=s_vec, post_div=s_vec, want_scales=False, out=out)
    q.ops.quant_group(w, 128, 4, False, no_clamp=True, want_scales=False, out=out)
    qw, qz, sc, _ = q.ops.quant_pack_awq(w, 128)
    q.ops.dequant_awq(qw, qz, sc, 128)
    q.ops.quant_rowwise(x, 8)
    q.ops.actquant_token_i8(x)
    q.ops.awq_wsum(w, 128)
    q.ops.sqdiff_sum(x, y)
    q.ops.colabsmax(x)
    q.ops.colstats(x, out_max=st_max, running=True, acc_maxsum=st_acc, acc_abssum=st_acc2)
torch.cuda.synchronize()
print("ok")
