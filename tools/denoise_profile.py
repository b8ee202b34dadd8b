"""Where a denoise step of the SD1.5 skeleton spends GPU time (torch profiler, 3 graph-free steps): python tools/denoise_profile.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
M = importlib.import_module("quantization---diffusion-models_b200.models")
dev = torch.device("cuda", 0)
model = M.StableDiffusion1_x.from_skeleton(device=dev)
model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq")
prompts = [f"p{i}" for i in range(8)]
lat = torch.randn(8, 4, 64, 64, generator=torch.Generator().manual_seed(1)).to(dev, model.pipeline.dtype)
model.generate(prompts, lat=lat, num_inference_steps=2)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.generate(prompts, lat=lat, num_inference_steps=3)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 3e3:.2f} ms per step")
for e in rows[:28]:
    print(f"{e.device_time_total / 3e3:8.3f} ms/step {100 * e.device_time_total / tot:5.1f}%  x{e.count // 3:4d}  {e.key[:110]}")
