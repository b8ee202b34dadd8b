"""Kernel breakdown of the AWQ search phases (scale + clip) on a 4-block SD3.5-L skeleton: python tools/calib_profile.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
M = importlib.import_module("quantization---diffusion-models_b200.models")
dev = torch.device("cuda", 0)
model = M.StableDiffusion3_5.from_skeleton(device=dev, layers=4)
model.calib_samples = model.default_calib_samples(4, 1)
model.calib_steps = 2
from torch.profiler import ProfilerActivity, profile
with torch.no_grad():
    model.pipeline(prompt=model.calib_samples[0][0], latents=model.calib_samples[0][1], num_inference_steps=1, guidance_scale=7.5)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.quantize(quant_config={"zero_point": True, "q_group_size": 128, "w_bit": 4, "version": "gemm"}, quantType="awq", calibrate=True)
    torch.cuda.synchronize()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in model.quantizer.timings.items()})
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.1f} ms")
for e in rows[:24]:
    print(f"{e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}%  x{e.count:5d}  {e.key[:120]}")
