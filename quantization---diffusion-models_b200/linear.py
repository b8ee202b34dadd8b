"""Real quantized Linear modules behind the reference's module contract.

WQLinear_GEMM  : the W4A16 module the reference imports from (un-vendored) AutoAWQ
                 (`awq.modules.linear.WQLinear_GEMM`; call sites quantize/quantizer.py:562-569,
                 models/base.py:1661-1663).  Same buffers (qweight [K, N/8] int32, qzeros [K/g, N/8] int32,
                 scales [K/g, N], bias [N]) and the same `from_linear(linear, w_bit, group_size, init_only,
                 scales, zeros)` signature; forward is the tcgen05 dequant-in-mainloop GEMM (kernel c).
W8A8Linear     : W8 per-channel x A8 per-token (quantize/fake_quant.py:86-93,109-118) with int8 codes and
                 the dequant scales in the GEMM epilogue (kernel d).  `smooth` carries a SmoothQuant
                 activation-side divisor when it was not folded into the previous op
                 (quantize/quantizer_SQ.py:425-431).
"""
import torch
from torch import nn

from . import ops


class WQLinear_GEMM(nn.Module):
    def __init__(self, w_bit, group_size, in_features, out_features, bias, dev, dtype=torch.float16):
        super().__init__()
        if w_bit != 4:
            raise NotImplementedError("Only 4-bit are supported for now.")
        self.in_features, self.out_features = in_features, out_features
        self.w_bit = w_bit
        self.repack = True   # keep a kernel-native copy of the weight for large-M calls (see _repacked)
        self.group_size = group_size if group_size != -1 else in_features
        assert self.in_features % self.group_size == 0
        assert out_features % (32 // self.w_bit) == 0
        self.register_buffer("qweight", torch.zeros((in_features, out_features // 8), dtype=torch.int32, device=dev))
        self.register_buffer("qzeros", torch.zeros((in_features // self.group_size, out_features // 8), dtype=torch.int32, device=dev))
        self.register_buffer("scales", torch.zeros((in_features // self.group_size, out_features), dtype=dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=dev))
        else:
            self.bias = None

    @classmethod
    def from_linear(cls, linear, w_bit, group_size, init_only=False, scales=None, zeros=None):
        """With scales/zeros=None the weight is quantised here by the fused RTN+pack kernel
        (pseudo_quantize_tensor + pack, quantizer.py:540-569 in one pass).  When the caller passes
        `scales`/`zeros` ([K/g, N], already transposed as quantizer.py:544-547 does) and a weight that
        is already fake-quantised, re-running the RTN reproduces exactly those codes, scales and zeros
        (quantisation of a fake-quantised tensor is idempotent); that is asserted, not assumed."""
        dev, dtype = linear.weight.device, linear.weight.dtype
        m = cls(w_bit, group_size, linear.in_features, linear.out_features, linear.bias is not None, dev, dtype)
        if init_only:
            return m
        qweight, qzeros, s, _ = ops.quant_pack_awq(linear.weight.data, m.group_size)
        if scales is not None and not torch.equal(s, scales.to(dtype)):
            raise ValueError("from_linear: the supplied scales do not match the weight's own RTN scales")
        m.qweight, m.qzeros, m.scales = qweight, qzeros, s
        if linear.bias is not None:
            m.bias = linear.bias.data.clone().to(dtype)
        return m

    def _repacked(self):
        """Kernel-native copy of the weight for the tensor-memory-A GEMM (ops.w4a16_repack_ts), built on first use and rebuilt
        when the buffers are replaced (load_state_dict, .to()).  Not part of the state dict: checkpoints keep the AWQ layout
        of utils/packing_utils.py only.  Costs 0.56 B / weight of device memory next to the 0.5 B / weight of qweight."""
        key = (self.qweight.data_ptr(), self.qzeros.data_ptr(), self.scales.data_ptr(), self.qweight._version)
        rp = self.__dict__.get("_rp")
        if rp is None or rp[0] != key:
            if self.group_size % 64 or self.in_features % 64:
                rp = (key, None)   # shapes the kernel does not tile
            else:
                rp = (key, ops.w4a16_repack_ts(self.qweight, self.qzeros, self.scales, self.group_size))
            self.__dict__["_rp"] = rp
        return rp[1]

    def forward(self, x):
        blob = self._repacked() if self.repack else None
        if x.dtype == self.scales.dtype:
            return ops.gemm_w4a16(x, self.qweight, self.qzeros, self.scales, self.group_size, self.bias, None, blob)
        return ops.gemm_w4a16(x.to(self.scales.dtype), self.qweight, self.qzeros, self.scales, self.group_size, self.bias, None, blob).to(x.dtype)

    def dequantize(self):
        """[N, K] fake-quant weight (utils/packing_utils.py:87-102, transposed back to nn.Linear layout)."""
        return ops.dequant_awq(self.qweight, self.qzeros, self.scales, self.group_size).t().contiguous()

    def extra_repr(self):
        return "in_features={}, out_features={}, bias={}, w_bit={}, group_size={}".format(
            self.in_features, self.out_features, self.bias is not None, self.w_bit, self.group_size)


class W8A8Linear(nn.Module):
    def __init__(self, in_features, out_features, bias, dev, dtype=torch.float16):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer("qweight", torch.zeros((out_features, in_features), dtype=torch.int8, device=dev))
        self.register_buffer("w_scales", torch.zeros(out_features, dtype=torch.float32, device=dev))
        self.register_buffer("smooth", None)
        if bias:
            self.register_buffer("bias", torch.zeros(out_features, dtype=dtype, device=dev))
        else:
            self.bias = None
        self.out_dtype = dtype

    @classmethod
    def from_float(cls, linear, smooth=None):
        dev, dtype = linear.weight.device, linear.weight.dtype
        m = cls(linear.in_features, linear.out_features, linear.bias is not None, dev, dtype)
        _, codes, scales, _ = ops.quant_rowwise(linear.weight.data, 8, want_dq=False, want_codes=True, want_scales=True)
        m.qweight, m.w_scales = codes, scales.float()
        if smooth is not None:
            m.smooth = smooth.to(device=dev, dtype=dtype)
        if linear.bias is not None:
            m.bias = linear.bias.data.clone().to(dtype)
        return m

    @torch.no_grad()
    def forward(self, x):
        xs = x if x.dtype == self.out_dtype else x.to(self.out_dtype)
        xq, sx = ops.actquant_token_i8(xs, self.smooth)
        y = ops.gemm_w8a8(xq, sx, self.qweight, self.w_scales, self.bias, out_dtype=self.out_dtype)
        y = y.reshape(*x.shape[:-1], self.out_features)
        return y if y.dtype == x.dtype else y.to(x.dtype)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, W8A8"


def w4a16_kernel_ok(in_features, out_features, group):
    """Shapes qdm_gemm_w4a16 tiles -- the ONE predicate the module swap and the kernel share (csrc/qdm_gemm.cu,
    gemm_w4a16_impl): K % 64 == 0, group = 64 * 2^j dividing K, N % 8 == 0.  Anything else (group_size = -1 on K = 320,
    group 192, ...) stays on the fake-quant path instead of building a module whose forward would raise."""
    j = group // 64 if group > 0 else 0
    return (group > 0 and group % 64 == 0 and (j & (j - 1)) == 0 and in_features % group == 0
            and in_features % 64 == 0 and out_features % 8 == 0)


def w8a8_kernel_ok(in_features, out_features):
    """qdm_gemm_w8a8: K % 16 == 0 (16-byte TMA rows of int8), N % 8 == 0."""
    return in_features % 16 == 0 and out_features % 8 == 0


def is_pointwise_conv(conv):
    """1x1, stride 1, no padding / dilation / groups: the convolution is `F.linear` over the channel dimension."""
    one = lambda v: tuple(v) == (1, 1) if not isinstance(v, int) else v == 1
    zero = lambda v: tuple(v) == (0, 0) if not isinstance(v, (int, str)) else v == 0
    return one(conv.kernel_size) and one(conv.stride) and zero(conv.padding) and one(conv.dilation) and conv.groups == 1


def nchw_as_tokens(x):
    """[B, C, H, W] -> [B*H*W, C] row-major.  A view when x is channels-last in memory (what this module's own
    output is, and what the transformer blocks after `proj_in` want anyway); one transpose copy otherwise."""
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b * h * w, c)


def tokens_as_nchw(y, b, h, w):
    """[B*H*W, O] -> logical [B, O, H, W] with channels-last strides (no copy): the tensor `F.conv2d` returns for a
    channels-last input."""
    return y.view(b, h, w, y.shape[-1]).permute(0, 3, 1, 2)


class QConv1x1(nn.Module):
    """a pointwise `nn.Conv2d` (SD1.5 `proj_in` / `proj_out`, resnet shortcuts) is a
    GEMM over channels, so its quantised replacement is the quantised Linear module on the [B*H*W, C] token view:
    W4A16 -> WQLinear_GEMM (kernel c), W8A8 -> W8A8Linear (kernel d).  The reference only fake-quantises conv
    weights and calls cuDNN (`WxAxConv2d`, fake_quant.py:263-398); 3x3 convolutions still do."""

    def __init__(self, inner, in_channels, out_channels):
        super().__init__()
        self.inner = inner
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = self.stride = self.dilation = (1, 1)
        self.padding = (0, 0)
        self.groups = 1

    @staticmethod
    def _as_linear(conv):
        lin = nn.Linear(conv.in_channels, conv.out_channels, bias=conv.bias is not None,
                        device="meta", dtype=conv.weight.dtype)
        lin.weight = nn.Parameter(conv.weight.data.reshape(conv.out_channels, conv.in_channels), requires_grad=False)
        if conv.bias is not None:
            lin.bias = nn.Parameter(conv.bias.data, requires_grad=False)
        return lin

    @classmethod
    def from_conv_w4a16(cls, conv, w_bit, group_size):
        if not is_pointwise_conv(conv):
            raise ValueError("QConv1x1 needs a 1x1 / stride 1 / ungrouped convolution")
        return cls(WQLinear_GEMM.from_linear(cls._as_linear(conv), w_bit, group_size), conv.in_channels, conv.out_channels)

    @classmethod
    def from_conv_w8a8(cls, conv, smooth=None):
        if not is_pointwise_conv(conv):
            raise ValueError("QConv1x1 needs a 1x1 / stride 1 / ungrouped convolution")
        return cls(W8A8Linear.from_float(cls._as_linear(conv), smooth), conv.in_channels, conv.out_channels)

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B, {self.in_channels}, H, W], got {tuple(x.shape)}")
        b, _, h, w = x.shape
        return tokens_as_nchw(self.inner(nchw_as_tokens(x)), b, h, w)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size=(1, 1)"


def is_conv3x3_gemm(conv):
    """3x3 / stride 1 or 2 / padding 1 / ungrouped with C % 64 == 0: the geometry `qdm_conv3x3_*` / `qdm_conv3x3s2_*` run
    as one implicit GEMM (stride 2 = the UNet down-samplers; grids the kernel cannot tile fall back to cuDNN at run time)."""
    t = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    return (t(conv.kernel_size) == (3, 3) and t(conv.stride) in ((1, 1), (2, 2)) and not isinstance(conv.padding, str)
            and t(conv.padding) == (1, 1) and t(conv.dilation) == (1, 1) and conv.groups == 1
            and conv.in_channels % 64 == 0 and conv.out_channels % 8 == 0)


def conv_group(k_flat, group_size):
    """quantisation group along the flattened (dy, dx, c) axis that the W4A16 kernel can tile: the reference's
    fallback loop (fake_quant.py:34-37) restricted to 64 * 2^j; 0 if none divides."""
    g = group_size if group_size > 0 else 128
    while g >= 64:
        if k_flat % g == 0 and g % 64 == 0 and ((g // 64) & (g // 64 - 1)) == 0:
            return g
        g -= 32
    return 0


class QConv3x3(nn.Module):
    """a 3x3 / padding 1 `nn.Conv2d` of stride 1 (the resnet convolutions, about half of
    the SD1.5 UNet's FLOPs) or stride 2 (the down-samplers, `qdm_conv3x3s2_nhwc_w4a16`) from PACKED int4 weights: the weight [N, C, 3, 3] is flattened tap-major to [N, 9C],
    quantised per group along that axis by the fused RTN + pack kernel (kernel b) and consumed by the W4A16 GEMM
    (kernel c) running as an implicit GEMM over the zero-padded NHWC grid (`qdm_conv3x3_w4a16`, no im2col buffer).
    Buffers follow WQLinear_GEMM (`qweight [9C, N/8]`, `qzeros`, `scales`, `bias`).  The reference only fake-quantises
    convolution weights and calls cuDNN (fake_quant.py:263-398)."""

    def __init__(self, w_bit, group_size, in_channels, out_channels, bias, dev, dtype=torch.float16, stride=1):
        super().__init__()
        if w_bit != 4:
            raise NotImplementedError("Only 4-bit are supported for now.")
        if stride not in (1, 2):
            raise ValueError("QConv3x3: stride must be 1 or 2")
        k = 9 * in_channels
        assert in_channels % 64 == 0 and out_channels % 8 == 0 and k % group_size == 0
        self.in_channels, self.out_channels, self.w_bit, self.group_size = in_channels, out_channels, w_bit, group_size
        self.kernel_size, self.stride, self.padding, self.dilation, self.groups = (3, 3), (stride, stride), (1, 1), (1, 1), 1
        self.register_buffer("qweight", torch.zeros((k, out_channels // 8), dtype=torch.int32, device=dev))
        self.register_buffer("qzeros", torch.zeros((k // group_size, out_channels // 8), dtype=torch.int32, device=dev))
        self.register_buffer("scales", torch.zeros((k // group_size, out_channels), dtype=dtype, device=dev))
        if bias:
            self.register_buffer("bias", torch.zeros(out_channels, dtype=dtype, device=dev))
        else:
            self.bias = None

    @classmethod
    def from_conv(cls, conv, w_bit, group_size, init_only=False):
        if not is_conv3x3_gemm(conv):
            raise ValueError("QConv3x3 needs a 3x3 / stride 1 or 2 / padding 1 / ungrouped convolution with C % 64 == 0")
        dev, dtype = conv.weight.device, conv.weight.dtype
        stride = conv.stride if isinstance(conv.stride, int) else conv.stride[0]
        m = cls(w_bit, group_size, conv.in_channels, conv.out_channels, conv.bias is not None, dev, dtype, stride)
        if init_only:
            return m
        m.qweight, m.qzeros, m.scales, _ = ops.quant_pack_awq(ops.conv3x3_weight_taps(conv.weight.data), group_size)
        if conv.bias is not None:
            m.bias = conv.bias.data.clone().to(dtype)
        return m

    @torch.no_grad()
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected [B, {self.in_channels}, H, W], got {tuple(x.shape)}")
        xs = x if x.dtype == self.scales.dtype else x.to(self.scales.dtype)
        st = self.stride[0]
        if st == 2 and not ops.conv3x3_stride2_ok(x.shape[2], x.shape[3]):
            # a grid the stride-2 tensor map cannot tile (odd sizes, output width not dividing 128): cuDNN on the
            # dequantised weight, as the reference does for every convolution (fake_quant.py:337-341)
            y = torch.nn.functional.conv2d(xs, self.dequantize().to(xs.dtype), self.bias, 2, 1)
        else:
            y = ops.conv3x3_w4a16(xs, self.qweight, self.qzeros, self.scales, self.group_size, self.bias, stride=st)
        return y if y.dtype == x.dtype else y.to(x.dtype)

    def dequantize(self):
        """[N, C, 3, 3] fake-quant weight (the conv layout back from the packed [9C, N/8] words)."""
        w = ops.dequant_awq(self.qweight, self.qzeros, self.scales, self.group_size).t()
        return w.reshape(self.out_channels, 3, 3, self.in_channels).permute(0, 3, 1, 2).contiguous()

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size=(3, 3), stride={self.stride}, padding=(1, 1), "
                f"w_bit={self.w_bit}, group_size={self.group_size}")
