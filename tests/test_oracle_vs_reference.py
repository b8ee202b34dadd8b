"""Live comparison of the oracle with the UNMODIFIED reference imported from /root/reference (tools/ref_shim.py), on
random inputs and sizes other than the committed fixtures.  Skipped where the reference tree is absent (the GPU box);
the committed fixtures (tests/test_oracle_golden.py) carry the same pinning there."""
import types

import numpy as np
import pytest
import torch

import ref_shim
import oracle.qdm_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not mounted on this box")
DT = {"f16": torch.float16, "bf16": torch.bfloat16}


@pytest.fixture(scope="module")
def ref():
    torch.set_grad_enabled(False)
    yield ref_shim.ref()
    torch.set_grad_enabled(True)


def eq(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    assert torch.equal(a, b), f"{what}: {(a != b).sum().item()} / {a.numel()} elements differ"


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pseudo_quantize_tensor_live(ref, dt, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    PQ = ref.quantizer.AwqQuantizer.pseudo_quantize_tensor
    for group, zp, bits in ((128, True, 4), (64, True, 4), (64, False, 4), (128, False, 8), (32, True, 3)):
        w = (torch.randn(40, 384, generator=g) * (0.02 + 0.5 * seed)).to(DT[dt])
        w[seed, :100] *= 30
        dq, s, z = PQ(types.SimpleNamespace(group_size=group, zero_point=zp), w.clone(), bitWidth=bits)
        odq, os_, oz, _ = O.rtn_group(w, group, zp, bits)
        eq(odq, dq, f"dq g{group} zp{zp} b{bits}")
        eq(os_, s, "scales")
        if zp:
            eq(oz, z, "zeros")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_fake_quant_functions_live(ref, dt):
    fq = ref.fake_quant
    g = torch.Generator().manual_seed(7)
    w = (torch.randn(48, 320, generator=g) * 0.05).to(DT[dt])
    eq(O.rtn_absmax_group(w, 4, 128)[0], fq.quantize_weight_absmax(w.clone(), n_bits=4, group_size=128, codeBookQuantInd=False), "group (K=320 -> 64)")
    eq(O.rtn_rows(w, 8)[0], fq.quantize_weight_per_channel_absmax(w, n_bits=8), "per channel")
    eq(O.rtn_tensor(w, 8)[0], fq.quantize_weight_per_tensor_absmax(w, n_bits=8), "per tensor")
    wc = (torch.randn(16, 8, 3, 3, generator=g) * 0.05).to(DT[dt])
    eq(O.rtn_rows(wc, 8)[0], fq.quantize_weight_per_channel_absmax(wc, n_bits=8), "conv per channel (rows of kw taps)")
    x = (torch.randn(3, 21, 320, generator=g) * 2).to(DT[dt])
    eq(O.rtn_rows(x.reshape(-1, 320), 8)[0].reshape(x.shape), fq.quantize_activation_per_token_absmax(x, n_bits=8), "per token")
    xc = torch.randn(2, 6, 8, 8, generator=g).to(DT[dt])
    eq(O.rtn_nchw_channel(xc, 8), fq.quantize_activation_per_channel_absmax(xc, n_bits=8), "nchw per channel")


def test_awq_layout_live(ref):
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 16, (192, 64), generator=g, dtype=torch.int32)        # [K, N]
    qu = ref.quant_utils
    packed = qu.pack(qu.apply_order(codes, "column", qu.AWQ_PACK_ORDER), "column")
    ours = O.awq_pack(codes.numpy())
    assert np.array_equal(ours, packed.numpy())
    assert np.array_equal(O.awq_unpack(ours), codes.numpy().astype(np.uint8))
    zeros = torch.randint(0, 16, (3, 64), generator=g, dtype=torch.int32)
    scales = (torch.rand(3, 64, generator=g) * 0.01 + 0.001).half()
    qz = O.awq_pack(zeros.numpy())
    want = ref.packing_utils.dequantize_gemm(torch.from_numpy(ours), torch.from_numpy(qz), scales, 4, 64)
    eq(O.awq_dequant(ours, qz, scales, 64), want, "dequantize_gemm")


@pytest.mark.parametrize("alpha", [0.5, 0.8])
def test_smooth_ln_fcs_live(ref, alpha):
    g = torch.Generator().manual_seed(11)
    C = 96
    ln = torch.nn.LayerNorm(C)
    ln.weight.data = (torch.rand(C, generator=g) + 0.5)
    ln.bias.data = torch.randn(C, generator=g) * 0.1
    fcs = [torch.nn.Linear(C, 64, bias=False) for _ in range(3)]
    for fc in fcs:
        fc.weight.data = torch.randn(64, C, generator=g) * 0.05
    ln, fcs = ln.half(), [fc.half() for fc in fcs]
    act = (torch.rand(C, generator=g) * 4 + 0.1).half()
    s = O.smooth_scales(act, [fc.weight.data for fc in fcs], alpha)
    lw, lb, ws = O.smooth_fold(ln.weight.data, ln.bias.data, [fc.weight.data for fc in fcs], s)
    SQ = ref.quantizer_SQ.SqQuantizer
    SQ.smooth_ln_fcs(SQ.__new__(SQ), ln, fcs, act, alpha=alpha)
    eq(lw, ln.weight.data, "ln.weight"), eq(lb, ln.bias.data, "ln.bias")
    for w, fc in zip(ws, fcs):
        eq(w, fc.weight.data, "fc.weight")


def test_wxax_conv_live(ref):
    g = torch.Generator().manual_seed(13)
    conv = torch.nn.Conv2d(64, 40, 3, padding=1)
    conv.weight.data = torch.randn(40, 64, 3, 3, generator=g) * 0.05
    conv.bias.data = torch.randn(40, generator=g)
    conv = conv.half()
    x = torch.randn(2, 64, 7, 9, generator=g).half()
    m = ref.fake_quant.WxAxConv2d.from_float(conv, weight_quant="per_tensor", act_quant="per_tensor", n_bits_W=8)
    wf = O.rtn_tensor(conv.weight.data, 8)[0]
    eq(wf, m.weight, "fake-quant conv weight")
    y, want = O.conv2d_fake(x, wf, conv.bias.data, 1, 1), m(x)
    assert ((y.float() - want.float()).abs().max() / want.float().abs().max()).item() <= 2e-3


# ------------------------------------------------------------------ host mirrors that are plain torch (no kernel): scale.py
class _Toy(torch.nn.Module):
    def __init__(self, g, dtype):
        super().__init__()
        C = 64
        self.ln = torch.nn.LayerNorm(C)
        self.q, self.k, self.v = (torch.nn.Linear(C, 48) for _ in range(3))
        self.fc1, self.fc2 = torch.nn.Linear(C, 96), torch.nn.Linear(96, C)
        for p in self.parameters():
            p.data = torch.randn(p.shape, generator=g) * 0.1 + (1.0 if p.dim() == 1 and p.numel() == C else 0.0)
        self.to(dtype)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_scale_and_clip_mirror_live(ref, dt):
    """apply_scale (LayerNorm -> q,k,v and fc1 -> fc2) and apply_clip of the product's scale.py against the reference's
    quantize/scale.py:25-153 on identical modules: every parameter bit-identical afterwards."""
    import importlib
    ours = importlib.import_module("quantization---diffusion-models_b200.scale")
    mods = []
    for _ in range(2):
        mods.append(_Toy(torch.Generator().manual_seed(5), DT[dt]))
    g = torch.Generator().manual_seed(6)
    s_ln = (torch.rand(64, generator=g) + 0.5).to(DT[dt])
    s_fc = (torch.rand(96, generator=g) + 0.5).to(DT[dt])
    clip = (torch.rand(48, 2, 1, generator=g) * 0.1 + 0.02).to(DT[dt])
    feats = [{"q": torch.randn(10, 64, generator=torch.Generator().manual_seed(7)).to(DT[dt])} for _ in range(2)]
    scales_list = lambda: [("ln", ("q", "k", "v"), s_ln.clone()), ("fc1", ("fc2",), s_fc.clone())]
    ref.scale.apply_scale(mods[0], scales_list(), input_feat_dict=feats[0])
    ours.apply_scale(mods[1], scales_list(), input_feat_dict=feats[1])
    ref.scale.apply_clip(mods[0], [("q", clip.clone())])
    ours.apply_clip(mods[1], [("q", clip.clone())])
    for (n, a), (_, b) in zip(mods[0].state_dict().items(), mods[1].state_dict().items()):
        eq(b, a, n)
    eq(feats[1]["q"], feats[0]["q"], "scaled input features")
