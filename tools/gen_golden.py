"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference via tools/ref_shim.py) on seeded inputs.  Run here (the reference does not travel
to the GPU box); the fixtures are committed and are what pins oracle/qdm_oracle.py and the CUDA path.

    python tools/gen_golden.py

Half / bfloat16 tensors are stored as their raw 16-bit patterns (key suffix `__f16` / `__bf16`).
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def enc(d, key, t):
    """store tensor `t` under `key` with a dtype-tag suffix"""
    if t is None:
        return
    t = t.detach().cpu().contiguous().clone()
    if t.dtype == torch.float16:
        d[key + "__f16"] = t.view(torch.int16).numpy()
    elif t.dtype == torch.bfloat16:
        d[key + "__bf16"] = t.view(torch.int16).numpy()
    else:
        d[key] = t.numpy()


def weight_like(shape, dtype, seed, outliers=True):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(shape, generator=g) * 0.05
    if outliers and len(shape) == 2:
        w[1, : shape[1] // 2] *= 40.0      # a wide-range group
        w[2, :] = 0.0                      # all-zero groups -> clamp(min=1e-5) path
        w[3, :] = 0.0123                   # constant groups (max == min)
        w[4, 5] = 3.0                      # single outlier
        w[5, :] = w[5, :].abs()            # all-positive group -> zero point 0
        w[6, :] = -w[6, :].abs()           # all-negative group -> zero point max
    return w.to(dtype)


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    r = ref_shim.ref()
    os.makedirs(OUT, exist_ok=True)
    PQ = r.quantizer.AwqQuantizer.pseudo_quantize_tensor

    # ---- A1 pseudo_quantize_tensor
    d = {}
    cases = []
    i = 0
    for dt in ("f16", "bf16", "f32"):
        for g in (128, 64, 32):
            for zp in (True, False):
                for bits in (4, 8, 3):
                    if dt == "f32" and g == 32:
                        continue
                    w = weight_like((24, 256), DT[dt], 100 + i)
                    self_ = types.SimpleNamespace(group_size=g, zero_point=zp)
                    dq, s, z = PQ(self_, w.clone(), bitWidth=bits)
                    tag = f"c{i}"
                    enc(d, tag + "_w", w), enc(d, tag + "_dq", dq), enc(d, tag + "_s", s), enc(d, tag + "_z", z)
                    cases.append(f"{tag},{dt},{g},{int(zp)},{bits}")
                    i += 1
    # group_size <= 0: per-row
    for dt in ("f16", "bf16"):
        w = weight_like((24, 200), DT[dt], 100 + i)
        self_ = types.SimpleNamespace(group_size=0, zero_point=True)
        dq, s, z = PQ(self_, w.clone(), bitWidth=4)
        tag = f"c{i}"
        enc(d, tag + "_w", w), enc(d, tag + "_dq", dq), enc(d, tag + "_s", s), enc(d, tag + "_z", z)
        cases.append(f"{tag},{dt},0,1,4")
        i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "pseudo_quantize_tensor.npz"), **d)

    # ---- A2..A6 fake_quant free functions
    fq = r.fake_quant
    d = {}
    cases = []
    i = 0
    for dt in ("f16", "bf16"):
        for k, g in ((256, 128), (320, 128), (192, 64)):
            for bits in (4, 8):
                w = weight_like((24, k), DT[dt], 300 + i)
                out = fq.quantize_weight_absmax(w.clone(), n_bits=bits, group_size=g, codeBookQuantInd=False)
                tag = f"g{i}"
                enc(d, tag + "_w", w), enc(d, tag + "_dq", out)
                cases.append(f"{tag},group,{dt},{g},{bits}")
                i += 1
        for shape in ((24, 320), (8, 6, 3, 3), (5, 7, 1, 1)):
            for bits in (8, 4):
                w = weight_like(shape, DT[dt], 300 + i)
                tag = f"g{i}"
                enc(d, tag + "_w", w), enc(d, tag + "_dq", fq.quantize_weight_per_channel_absmax(w, n_bits=bits))
                cases.append(f"{tag},channel,{dt},0,{bits}")
                i += 1
        w = weight_like((24, 320), DT[dt], 300 + i)
        tag = f"g{i}"
        enc(d, tag + "_w", w), enc(d, tag + "_dq", fq.quantize_weight_per_tensor_absmax(w, n_bits=8))
        cases.append(f"{tag},tensor,{dt},0,8")
        i += 1
        x = (torch.randn(2, 19, 320, generator=torch.Generator().manual_seed(300 + i)) * 2).to(DT[dt])
        x[0, 3, 7] = 60.0
        tag = f"g{i}"
        enc(d, tag + "_w", x), enc(d, tag + "_dq", fq.quantize_activation_per_token_absmax(x, n_bits=8))
        cases.append(f"{tag},token,{dt},0,8")
        i += 1
        x = (torch.randn(2, 6, 8, 8, generator=torch.Generator().manual_seed(300 + i))).to(DT[dt])
        tag = f"g{i}"
        enc(d, tag + "_w", x), enc(d, tag + "_dq", fq.quantize_activation_per_channel_absmax(x, n_bits=8))
        cases.append(f"{tag},nchw,{dt},0,8")
        i += 1
        x = (torch.randn(3, 11, 64, generator=torch.Generator().manual_seed(300 + i))).to(DT[dt])
        tag = f"g{i}"
        enc(d, tag + "_w", x), enc(d, tag + "_dq", fq.quantize_activation_per_tensor_absmax(x, n_bits=8))
        cases.append(f"{tag},tensor,{dt},0,8")
        i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "fake_quant.npz"), **d)

    # ---- A15 layout: reference pack <-> dequantize_gemm
    d = {}
    g = torch.Generator().manual_seed(7)
    K, N, gs = 128, 64, 64
    codes = torch.randint(0, 16, (K, N), generator=g, dtype=torch.int32)
    zeros = torch.randint(0, 16, (K // gs, N), generator=g, dtype=torch.int32)
    scales = (torch.rand(K // gs, N, generator=g) * 0.02 + 0.001).to(torch.float16)
    qu, pu = r.quant_utils, r.packing_utils
    qweight = qu.pack(qu.apply_order(codes, "column", qu.AWQ_PACK_ORDER), "column")
    qzeros = qu.pack(qu.apply_order(zeros, "column", qu.AWQ_PACK_ORDER), "column")
    deq = pu.dequantize_gemm(qweight, qzeros, scales, 4, gs)
    iw, iz = pu.unpack_awq(qweight, qzeros, 4)
    iw, iz = pu.reverse_awq_order(iw, iz, 4)
    assert torch.equal(iw & 0xF, codes.to(torch.int8))
    d["codes"], d["zeros"] = codes.numpy(), zeros.numpy()
    d["qweight"], d["qzeros"] = qweight.numpy(), qzeros.numpy()
    enc(d, "scales", scales), enc(d, "deq", deq)
    d["group"] = np.array(gs)
    np.savez_compressed(os.path.join(OUT, "awq_layout.npz"), **d)

    # ---- A9/A10/A11 scale + clip search on a toy LayerNorm -> q,k,v group
    d = {}
    cases = []
    i = 0
    Q = r.quantizer.AwqQuantizer
    for dt in ("f16", "bf16"):
        for zp in (True, False):
            C, T = 128, 96
            gen = torch.Generator().manual_seed(500 + i)
            x = torch.randn(2, T, C, generator=gen)
            x[..., 5] *= 30.0
            x[..., 77] *= 12.0
            x = x.to(DT[dt])
            lins = [torch.nn.Linear(C, 64, bias=True) for _ in range(3)]
            for li, lin in enumerate(lins):
                lin.weight.data = weight_like((64, C), torch.float32, 600 + 10 * i + li, outliers=False)
                lin.bias.data = torch.randn(64, generator=gen) * 0.01
                lin.to(DT[dt])

            class Cat(torch.nn.Module):
                def __init__(self, ls):
                    super().__init__()
                    self.ls = torch.nn.ModuleList(ls)

                def forward(self, x):
                    return torch.cat([l(x) for l in self.ls], dim=-1)

                def state_dict(self, *a, **k):
                    # quantizer.py:703 snapshots `v.cpu()` -- a real copy on the reference's CUDA path, but
                    # an alias on a CPU-resident module (the in-place mul_ at :727 would then corrupt the
                    # snapshot).  Returning clones reproduces the CUDA-path semantics on this CPU box.
                    return {n: v.clone() for n, v in super().state_dict(*a, **k).items()}

            block = Cat(lins)
            captured = {}
            self_ = types.SimpleNamespace(group_size=64, zero_point=zp, duo_scaling=True, max_chunk_memory=1 << 30,
                                          n_parallel_calib_samples=None)
            self_.pseudo_quantize_tensor = types.MethodType(Q.pseudo_quantize_tensor, self_)
            self_._module_forward = types.MethodType(Q._module_forward, self_)
            self_._compute_loss = types.MethodType(Q._compute_loss, self_)
            self_._sanitize_kwargs = lambda kw, m: {}
            orig = Q._compute_best_scale

            def spy(self, x_, w_mean, x_mean, *a, **k):
                captured["w_mean"], captured["x_mean"] = w_mean.clone(), x_mean.clone()
                return orig(self, x_, w_mean, x_mean, *a, **k)

            self_._compute_best_scale = types.MethodType(spy, self_)
            w_before = [l.weight.data.clone() for l in lins]
            _, _, best = Q._search_best_scale(self_, block, lins[0], lins, x, module2inspect=block, kwargs={})
            for l, wb in zip(lins, w_before):
                assert torch.equal(l.weight.data, wb)
            clip = Q._compute_best_clip(self_, lins[0].weight.data, x)
            tag = f"s{i}"
            enc(d, tag + "_x", x)
            for li, lin in enumerate(lins):
                enc(d, f"{tag}_w{li}", lin.weight.data), enc(d, f"{tag}_b{li}", lin.bias.data)
            enc(d, tag + "_wmean", captured["w_mean"]), enc(d, tag + "_xmean", captured["x_mean"])
            enc(d, tag + "_best", best), enc(d, tag + "_clip", clip)
            cases.append(f"{tag},{dt},64,{int(zp)}")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "awq_search.npz"), **d)

    # ---- A13/A14 SmoothQuant
    d = {}
    cases = []
    i = 0
    SQ = r.quantizer_SQ.SqQuantizer
    for dt in ("f16", "bf16"):
        for alpha in (0.5, 0.8):
            C = 96
            gen = torch.Generator().manual_seed(700 + i)
            ln = torch.nn.LayerNorm(C)
            ln.weight.data = 1 + 0.1 * torch.randn(C, generator=gen)
            ln.bias.data = 0.1 * torch.randn(C, generator=gen)
            ln.to(DT[dt])
            fcs = [torch.nn.Linear(C, 48, bias=False).to(DT[dt]) for _ in range(3)]
            for li, fc in enumerate(fcs):
                fc.weight.data = weight_like((48, C), DT[dt], 800 + 10 * i + li, outliers=False)
            hook = r.calib_data.Mean_Max_Activation_Hook()
            calls = []
            for c in range(3):
                xin = torch.randn(2, 17, C, generator=gen)
                xin[..., 11] *= 25
                xin = xin.to(DT[dt])
                calls.append(xin)
                hook(None, (xin,), None)
            per_call = [hook.max_scales[k] for k in sorted(hook.max_scales)]
            act = torch.stack(per_call).mean(dim=0)
            tag = f"q{i}"
            enc(d, tag + "_lnw", ln.weight.data), enc(d, tag + "_lnb", ln.bias.data)
            for li, fc in enumerate(fcs):
                enc(d, f"{tag}_w{li}", fc.weight.data)
            for c, xin in enumerate(calls):
                enc(d, f"{tag}_x{c}", xin), enc(d, f"{tag}_max{c}", per_call[c])
            enc(d, tag + "_act", act)
            SQ.smooth_ln_fcs(None, ln, fcs, act.clone(), alpha=alpha)
            enc(d, tag + "_lnw_out", ln.weight.data), enc(d, tag + "_lnb_out", ln.bias.data)
            for li, fc in enumerate(fcs):
                enc(d, f"{tag}_w{li}_out", fc.weight.data)
            cases.append(f"{tag},{dt},{alpha}")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "smoothquant.npz"), **d)

    # ---- A7 WxAxLinear.from_float + forward
    d = {}
    cases = []
    i = 0
    for dt in ("f16",):
        for wq, bits, g in (("group", 4, 128), ("per_channel", 8, 0), ("per_tensor", 8, 0)):
            lin = torch.nn.Linear(256, 48, bias=True)
            lin.weight.data = weight_like((48, 256), torch.float32, 900 + i, outliers=False)
            lin.to(DT[dt])
            x = torch.randn(3, 10, 256, generator=torch.Generator().manual_seed(950 + i)).to(DT[dt])
            m = fq.WxAxLinear.from_float(lin, weight_quant=wq, n_bits_W=bits, group_size_W=g, codeBookQuantInd=False)
            y = m(x)
            tag = f"l{i}"
            enc(d, tag + "_w", lin.weight.data), enc(d, tag + "_b", lin.bias.data), enc(d, tag + "_x", x)
            enc(d, tag + "_wq", m.weight), enc(d, tag + "_y", y)
            cases.append(f"{tag},{dt},{wq},{bits},{g}")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "wxax_linear.npz"), **d)
    gen_wxax_conv(r)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def gen_wxax_conv(r):
    """A8 WxAxConv2d.from_float + forward (fake_quant.py:263-398) for the geometries the B200 path runs as GEMMs:
    1x1 (pointwise) and 3x3 / stride 1 / padding 1, per_tensor and per_channel weights, fp16, reference on CPU."""
    fq = r.fake_quant
    d, cases, i = {}, [], 0
    for ksz, cin, cout, hw in ((1, 64, 72, 6), (3, 64, 64, 8), (3, 128, 72, 5)):
        for wq, bits in (("per_tensor", 8), ("per_channel", 8)):
            g = torch.Generator().manual_seed(1200 + i)
            conv = torch.nn.Conv2d(cin, cout, ksz, padding=ksz // 2, bias=True)
            conv.weight.data = (torch.randn(cout, cin, ksz, ksz, generator=g) * 0.05)
            conv.bias.data = torch.randn(cout, generator=g)
            conv.to(torch.float16)
            x = torch.randn(2, cin, hw, hw + (1 if ksz == 3 and hw == 5 else 0), generator=g).to(torch.float16)
            m = fq.WxAxConv2d.from_float(conv, weight_quant=wq, act_quant="per_tensor", n_bits_W=bits)
            y = m(x)
            tag = f"c{i}"
            enc(d, tag + "_w", conv.weight.data), enc(d, tag + "_b", conv.bias.data), enc(d, tag + "_x", x)
            enc(d, tag + "_wq", m.weight), enc(d, tag + "_y", y)
            cases.append(f"{tag},f16,{wq},{bits},{ksz}")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "wxax_conv.npz"), **d)


def gen_wxax_conv_s2(r):
    """The same for the UNet down-samplers (Downsample2D: 3x3, stride 2, padding 1), which the B200 path runs as an
    implicit GEMM behind a tensor map with element strides of 2 (qdm_conv3x3s2_nhwc_*); reference on CPU, fp16."""
    fq = r.fake_quant
    d, cases, i = {}, [], 0
    for cin, cout, h, w, wqs in ((64, 64, 8, 8, ("per_tensor", "per_channel")), (128, 72, 16, 16, ("per_channel",)),
                                 (64, 72, 15, 16, ("per_tensor",))):   # the last: an odd grid (stays on cuDNN)
        for wq, bits in ((q_, 8) for q_ in wqs):
            g = torch.Generator().manual_seed(1400 + i)
            conv = torch.nn.Conv2d(cin, cout, 3, stride=2, padding=1, bias=True)
            conv.weight.data = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
            conv.bias.data = torch.randn(cout, generator=g)
            conv.to(torch.float16)
            x = torch.randn(2, cin, h, w, generator=g).to(torch.float16)
            m = fq.WxAxConv2d.from_float(conv, weight_quant=wq, act_quant="per_tensor", n_bits_W=bits)
            y = m(x)
            tag = f"s{i}"
            enc(d, tag + "_w", conv.weight.data), enc(d, tag + "_b", conv.bias.data), enc(d, tag + "_x", x)
            enc(d, tag + "_wq", m.weight), enc(d, tag + "_y", y)
            cases.append(f"{tag},f16,{wq},{bits},3")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "wxax_conv_s2.npz"), **d)


def gen_act_quant(r):
    """A6 / A5 activation fake quantisers that the first fixtures did not cover: the NCHW per-(n, c, patch) quantiser
    (fake_quant.py:134-153, incl. its `group_size -= 2` fallback) and the 16-bit forms the reference's default
    a_bit = 16 reaches (q_max = 32767, fake_quant.py:109-118,158-167)."""
    import contextlib
    import io
    fq = r.fake_quant
    d, cases, i = {}, [], 0
    for dt in ("f16", "bf16"):
        for shape, gs in (((2, 6, 16, 16), 8), ((1, 5, 12, 12), 4), ((2, 3, 16, 16), 128), ((1, 4, 24, 24), 10)):
            g = torch.Generator().manual_seed(1300 + i)
            x = torch.randn(shape, generator=g)
            x[:, 1] *= 25.0
            x[0, 0, :4, :4] = 0.0                       # an all-zero patch -> clamp(min=1e-5) path
            x = x.to(DT[dt])
            with contextlib.redirect_stdout(io.StringIO()):   # the reference prints the group size it settled on
                y = fq.quantize_activation_per_channel_group_absmax(x, group_size=gs, n_bits=8)
            tag = f"g{i}"
            enc(d, tag + "_x", x), enc(d, tag + "_y", y)
            cases.append(f"{tag},patch,{dt},{gs},8")
            i += 1
        for kind, bits in (("token", 16), ("token", 12), ("tensor", 16), ("nchw", 16)):
            g = torch.Generator().manual_seed(1300 + i)
            x = torch.randn((3, 7, 40) if kind != "nchw" else (2, 5, 6, 6), generator=g) * 3.0
            x[..., 3] *= 30.0
            x = x.to(DT[dt])
            fn = {"token": fq.quantize_activation_per_token_absmax, "tensor": fq.quantize_activation_per_tensor_absmax,
                  "nchw": fq.quantize_activation_per_channel_absmax}[kind]
            tag = f"g{i}"
            enc(d, tag + "_x", x), enc(d, tag + "_y", fn(x, n_bits=bits))
            cases.append(f"{tag},{kind},{dt},0,{bits}")
            i += 1
    d["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "act_quant.npz"), **d)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "conv":   # only the convolution fixture (added later; the others are unchanged)
        torch.set_grad_enabled(False)
        gen_wxax_conv(ref_shim.ref())
    elif len(sys.argv) > 1 and sys.argv[1] == "conv_s2":   # only the stride-2 convolution fixture (round 2, later)
        torch.set_grad_enabled(False)
        gen_wxax_conv_s2(ref_shim.ref())
    elif len(sys.argv) > 1 and sys.argv[1] == "acts":  # only the activation-quantiser fixture (round 2)
        torch.set_grad_enabled(False)
        gen_act_quant(ref_shim.ref())
    else:
        main()
