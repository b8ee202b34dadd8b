"""CPU: the oracle restatement reproduces the reference's own outputs (tests/golden/*.npz were produced
by tools/gen_golden.py from the unmodified reference).  Bit-exact for every floating-point chain."""
import numpy as np
import pytest
import torch

from _util import DT, Golden, assert_bit_equal

import oracle.qdm_oracle as O


def test_pseudo_quantize_tensor_matches_reference():
    g = Golden("pseudo_quantize_tensor.npz")
    assert len(g.cases()) >= 40
    for tag, dt, group, zp, bits in g.cases():
        w = g.get(tag + "_w")
        dq, s, z, codes = O.rtn_group(w, int(group), bool(int(zp)), int(bits))
        assert_bit_equal(dq, g.get(tag + "_dq"), f"{tag} dq")
        assert_bit_equal(s, g.get(tag + "_s"), f"{tag} scales")
        if int(zp):
            assert_bit_equal(z, g.get(tag + "_z"), f"{tag} zeros")
            assert codes.min() >= 0 and codes.max() <= 2 ** int(bits) - 1
        else:
            assert g.get(tag + "_z") is None
            assert codes.min() >= -(2 ** (int(bits) - 1)) and codes.max() <= 2 ** (int(bits) - 1) - 1


def test_fake_quant_functions_match_reference():
    g = Golden("fake_quant.npz")
    kinds = set()
    for tag, kind, dt, group, bits in g.cases():
        w, want = g.get(tag + "_w"), g.get(tag + "_dq")
        kinds.add(kind)
        if kind == "group":
            got = O.rtn_absmax_group(w, int(bits), int(group))[0]
        elif kind in ("channel", "token"):
            got = O.rtn_rows(w, int(bits))[0]
        elif kind == "tensor":
            got = O.rtn_tensor(w, int(bits))[0]
        elif kind == "nchw":
            got = O.rtn_nchw_channel(w, int(bits))
        assert_bit_equal(got, want, f"{tag} {kind}")
    assert kinds == {"group", "channel", "token", "tensor", "nchw"}


def test_activation_quantisers_round2_fixture():
    """act_quant.npz (tools/gen_golden.py acts): the NCHW per-patch quantiser (fake_quant.py:134-153) and the 16-bit
    forms of the per-token / per-tensor / per-(n, c) quantisers the reference's default a_bit = 16 reaches."""
    g = Golden("act_quant.npz")
    kinds = set()
    for tag, kind, dt, gs, bits in g.cases():
        x, want = g.get(tag + "_x"), g.get(tag + "_y")
        kinds.add(kind)
        got = {"patch": lambda: O.rtn_nchw_patch(x, int(gs), int(bits)), "token": lambda: O.rtn_rows(x, int(bits))[0],
               "tensor": lambda: O.rtn_tensor(x, int(bits))[0], "nchw": lambda: O.rtn_nchw_channel(x, int(bits))}[kind]()
        assert_bit_equal(got, want, f"{tag} {kind} {dt}")
    assert kinds == {"patch", "token", "tensor", "nchw"}


def test_group_fallback():
    assert O.effective_group(320, 128) == 64      # fake_quant.py:34-37
    assert O.effective_group(2432, 128) == 128
    assert O.effective_group(192, 128) == 96


def test_awq_layout_matches_reference():
    g = Golden("awq_layout.npz")
    codes, zeros = g.get("codes").numpy(), g.get("zeros").numpy()
    assert np.array_equal(O.awq_pack(codes), g.get("qweight").numpy())
    assert np.array_equal(O.awq_pack(zeros), g.get("qzeros").numpy())
    assert np.array_equal(O.awq_unpack(g.get("qweight").numpy()), codes.astype(np.uint8))
    deq = O.awq_dequant(g.get("qweight").numpy(), g.get("qzeros").numpy(), g.get("scales"), int(g.get("group")))
    assert_bit_equal(deq, g.get("deq"), "dequantize_gemm")


def test_awq_layout_known_answer():
    # nibble i of word c holds column 8c + AWQ_ORDER[i]  (utils/packing_utils.py:4)
    codes = np.arange(16, dtype=np.int32).reshape(1, 16) % 16
    q = O.awq_pack(codes).view(np.uint32)
    assert q[0, 0] == sum(O.AWQ_ORDER[i] << (4 * i) for i in range(8))
    assert q[0, 1] == sum((8 + O.AWQ_ORDER[i]) << (4 * i) for i in range(8))


def _toy_forward(x, biases):
    def fwd(ws):
        return torch.cat([torch.nn.functional.linear(x, w, b) for w, b in zip(ws, biases)], dim=-1)
    return fwd


def test_awq_search_matches_reference():
    g = Golden("awq_search.npz")
    for tag, dt, group, zp in g.cases():
        x = g.get(tag + "_x")
        ws = [g.get(f"{tag}_w{i}") for i in range(3)]
        bs = [g.get(f"{tag}_b{i}") for i in range(3)]
        assert_bit_equal(O.awq_w_mean(ws, int(group)), g.get(tag + "_wmean"), f"{tag} w_mean")
        assert_bit_equal(O.awq_x_mean(x), g.get(tag + "_xmean"), f"{tag} x_mean")
        best, ratio, hist = O.awq_search_scale(x, ws, _toy_forward(x, bs), int(group), bool(int(zp)), 4, True)
        assert_bit_equal(best, g.get(tag + "_best"), f"{tag} best_scales")
        assert len(hist) == 20 and 0 <= ratio < 1
        clip = O.awq_search_clip(ws[0], x, int(group), bool(int(zp)), 4)
        assert_bit_equal(clip, g.get(tag + "_clip"), f"{tag} best clip")


def test_smoothquant_matches_reference():
    g = Golden("smoothquant.npz")
    for tag, dt, alpha in g.cases():
        per_call = []
        for c in range(3):
            m = O.hook_colabsmax(g.get(f"{tag}_x{c}"))
            assert_bit_equal(m, g.get(f"{tag}_max{c}"), f"{tag} hook {c}")
            per_call.append(m)
        act = O.mean_of_calls(per_call)
        assert_bit_equal(act, g.get(tag + "_act"), f"{tag} act")
        ws = [g.get(f"{tag}_w{i}") for i in range(3)]
        s = O.smooth_scales(act, ws, float(alpha))
        lnw, lnb, ws2 = O.smooth_fold(g.get(tag + "_lnw"), g.get(tag + "_lnb"), ws, s)
        assert_bit_equal(lnw, g.get(tag + "_lnw_out"), f"{tag} ln.weight")
        assert_bit_equal(lnb, g.get(tag + "_lnb_out"), f"{tag} ln.bias")
        for i in range(3):
            assert_bit_equal(ws2[i], g.get(f"{tag}_w{i}_out"), f"{tag} fc{i}.weight")


def test_wxax_linear_matches_reference():
    g = Golden("wxax_linear.npz")
    for tag, dt, wq, bits, group in g.cases():
        w, b, x = g.get(tag + "_w"), g.get(tag + "_b"), g.get(tag + "_x")
        if wq == "group":
            wf = O.rtn_absmax_group(w, int(bits), int(group))[0]
        elif wq == "per_channel":
            wf = O.rtn_rows(w, int(bits))[0]
        else:
            wf = O.rtn_tensor(w, int(bits))[0]
        assert_bit_equal(wf, g.get(tag + "_wq"), f"{tag} fake-quant weight")
        y = O.linear_fake(x, wf, b)
        ref = g.get(tag + "_y")
        # F.linear accumulates in a backend-dependent order: tolerance, not bit equality
        assert ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item() <= 2e-3


def test_wxax_conv_matches_reference():
    """A8: WxAxConv2d.from_float + forward of the reference (fake_quant.py:263-398), 1x1 and 3x3 / pad 1, vs the oracle."""
    g = Golden("wxax_conv.npz")
    for tag, dt, wq, bits, ksz in g.cases():
        w, b, x = g.get(tag + "_w"), g.get(tag + "_b"), g.get(tag + "_x")
        wf = (O.rtn_rows(w, int(bits))[0] if wq == "per_channel" else O.rtn_tensor(w, int(bits))[0]).reshape(w.shape)
        assert_bit_equal(wf, g.get(tag + "_wq"), f"{tag} fake-quant conv weight")
        y = O.conv2d_fake(x, wf, b, 1, int(ksz) // 2)
        ref = g.get(tag + "_y")
        assert y.shape == ref.shape
        assert ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item() <= 2e-3


def test_wxax_conv_stride2_matches_reference():
    """A8 for the UNet down-samplers (3x3, stride 2, padding 1): the reference's WxAxConv2d.from_float + forward
    (fake_quant.py:263-398 keeps the module's stride) vs the oracle; fixture from tools/gen_golden.py conv_s2."""
    g = Golden("wxax_conv_s2.npz")
    n = 0
    for tag, dt, wq, bits, ksz in g.cases():
        w, b, x = g.get(tag + "_w"), g.get(tag + "_b"), g.get(tag + "_x")
        wf = (O.rtn_rows(w, int(bits))[0] if wq == "per_channel" else O.rtn_tensor(w, int(bits))[0]).reshape(w.shape)
        assert_bit_equal(wf, g.get(tag + "_wq"), f"{tag} fake-quant conv weight")
        y = O.conv2d_fake(x, wf, b, 2, 1)
        ref = g.get(tag + "_y")
        assert y.shape == ref.shape == (x.shape[0], w.shape[0], (x.shape[2] + 1) // 2, (x.shape[3] + 1) // 2)
        assert ((y.float() - ref.float()).abs().max() / ref.float().abs().max()).item() <= 2e-3
        n += 1
    assert n == 4
