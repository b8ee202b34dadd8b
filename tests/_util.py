"""Shared test helpers: golden-fixture decoding and tensor comparison."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


class Golden:
    """npz written by tools/gen_golden.py; 16-bit floats are stored as raw bit patterns."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    def cases(self):
        return [str(c).split(",") for c in self.z["cases"]]

    def has(self, key):
        return any(k in self.z for k in (key, key + "__f16", key + "__bf16"))

    def get(self, key):
        if key + "__f16" in self.z:
            return torch.from_numpy(self.z[key + "__f16"].copy()).view(torch.float16)
        if key + "__bf16" in self.z:
            return torch.from_numpy(self.z[key + "__bf16"].copy()).view(torch.bfloat16)
        if key in self.z:
            return torch.from_numpy(self.z[key].copy())
        return None


def assert_bit_equal(a, b, what=""):
    """value equality element by element (+0.0 == -0.0), same dtype and shape"""
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a, b = a.cpu(), b.cpu()
    if not torch.equal(a, b):
        bad = (a != b).nonzero()
        i = tuple(bad[0].tolist())
        raise AssertionError(f"{what}: {bad.shape[0]} / {a.numel()} mismatches, first at {i}: {a[i].item()} vs {b[i].item()}")


def max_rel_err(y, ref):
    """max |y - ref| / max |ref| -- the north-star GEMM tolerance metric (<= 1e-2)."""
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()
