"""Host-side mirror of the reference's quantize/fake_quant.py: same function and class names, argument
meaning and error behaviour; every arithmetic step runs in libqdm's sm_100a kernels.

Deviations from the reference, all deliberate:
  * nothing mutates its input in place (the reference does when `reshape` returns a view,
    fake_quant.py:40,57,72);
  * `codeBookQuantInd=True` (experimental k-means codebooks, off by default at models/base.py:380-385)
    raises NotImplementedError;
  * WxAxLinear.forward does not evaluate `str(self)` per call (fake_quant.py:216-217).
"""
import os
from functools import partial

import torch
from torch import nn

from . import ops


def _effective_group(k, group_size):
    """fake_quant.py:34-37: shrink the group by 32 until it divides K (K=320 -> 64)."""
    while group_size > 0 and k % group_size != 0:
        group_size -= 32
    if group_size < 0:
        raise ZeroDivisionError("group size fell below zero while searching for a divisor")
    return group_size


def quantize_weight_absmax(w, n_bits=8, group_size=0, codeBookQuantInd=False, debugPath=[], debug=False):
    """fake_quant.py:21-84: symmetric per-group absmax RTN, codes not clamped, result fp16."""
    if codeBookQuantInd:
        raise NotImplementedError("codebook (k-means) quantisation is outside the quantized-linear hot path")
    shape = w.shape
    if group_size > 0:
        group_size = _effective_group(shape[-1], group_size)
        if group_size == 0:
            raise ZeroDivisionError("integer modulo by zero")  # what the reference hits for 4-D weights
    else:
        assert w.dim() == 2
    dq = ops.quant_group(w, group_size if group_size > 0 else shape[-1], n_bits, zero_point=False, no_clamp=True,
                         want_scales=False)[0]
    return dq.reshape(shape).to(torch.float16)


def quantize_weight_per_channel_absmax(w, n_bits=8):
    """fake_quant.py:86-93: one scale per last-dim row (conv weights: per (o, i, kh) row of kw taps)."""
    return ops.quant_rowwise(w, n_bits)[0].reshape(w.shape)


@torch.no_grad()
def quantize_weight_per_tensor_absmax(w, n_bits=8):
    """fake_quant.py:97-105."""
    return ops.quant_tensor(w, n_bits)[0].reshape(w.shape)


@torch.no_grad()
def quantize_activation_per_token_absmax(t, n_bits=8):
    """fake_quant.py:109-118."""
    return ops.quant_rowwise(t, n_bits)[0].reshape(t.shape)


@torch.no_grad()
def quantize_activation_per_channel_absmax(t, n_bits=8):
    """fake_quant.py:124-131: NCHW, one scale per (n, c) over H x W."""
    n, c, h, w = t.shape
    return ops.quant_rowwise(t.contiguous().reshape(n * c, h * w), n_bits)[0].reshape(t.shape)


@torch.no_grad()
def quantize_activation_per_channel_group_absmax(t, group_size=128, n_bits=8):
    """fake_quant.py:134-153: one scale per (n, c, patch) with square spatial patches."""
    N, C, H, W = t.shape
    while H % group_size != 0 or W % group_size != 0:
        group_size -= 2
    g = group_size
    patches = t.reshape(N, C, H // g, g, W // g, g).permute(0, 1, 2, 4, 3, 5).contiguous()  # [..., g, g] rows
    q = ops.quant_rowwise(patches.reshape(-1, g * g), n_bits)[0]
    return q.reshape(N, C, H // g, W // g, g, g).permute(0, 1, 2, 4, 3, 5).reshape(N, C, H, W)


@torch.no_grad()
def quantize_activation_per_tensor_absmax(t, n_bits=8):
    """fake_quant.py:158-167."""
    return ops.quant_tensor(t, n_bits)[0].reshape(t.shape)


class WxAxLinear(nn.Module):
    """fake_quant.py:170-261: the nn.Linear replacement holding fake-quantised fp16 weights.
    forward = F.linear on the tcgen05 GEMM (kernel c/d family, qdm_gemm_f16)."""

    def __init__(self, in_features, out_features, bias=True, weight_quant='per_channel', act_quant='per_token',
                 quantize_output=False, n_bits_A=16, q_act=False, device=None):
        """`device` (new): where the buffers are created; the reference fills them with CPU randn first
        (fake_quant.py:179-183), which costs seconds on a UNet and is overwritten by from_float anyway."""
        super().__init__()
        self.scales = []
        self.inputs = []
        self.quantize_act = q_act
        self.in_features = in_features
        self.out_features = out_features
        self.n_bits_A, self.quantize_output = n_bits_A, quantize_output
        self.register_buffer('weight', torch.empty(out_features, in_features, dtype=torch.float16, device=device))
        if bias:
            self.register_buffer('bias', torch.zeros(out_features, dtype=torch.float16, device=device))
        else:
            self.register_buffer('bias', None)
        self.weight_quant_name = weight_quant
        if act_quant == 'per_token':
            self.act_quant_name = 'per_token'
            self.act_quant = partial(quantize_activation_per_token_absmax, n_bits=n_bits_A)
        elif act_quant == 'per_tensor':
            self.act_quant_name = 'per_tensor'
            self.act_quant = partial(quantize_activation_per_tensor_absmax, n_bits=n_bits_A)
        else:
            raise ValueError(f'Invalid act_quant: {act_quant}')
        if quantize_output:
            self.output_quant_name = self.act_quant_name
            self.output_quant = self.act_quant
        else:
            self.output_quant_name = 'None'
            self.output_quant = lambda x: x

    @torch.no_grad()
    def forward(self, x):
        q_x = self.act_quant(x) if self.quantize_act else x
        w = self.weight if self.weight.dtype == q_x.dtype else self.weight.to(q_x.dtype)
        y = ops.gemm_f16(q_x, w, self.bias)
        return self.output_quant(y).to(x.dtype)

    @classmethod
    def from_linear(cls, module, init_only=False, weight_quant='per_channel', act_quant='per_token',
                    quantize_output=False, n_bits_W=8, n_bits_A=16, group_size_W=0):
        assert isinstance(module, torch.nn.Linear)
        return cls(module.in_features, module.out_features, module.bias is not None, act_quant=act_quant,
                   quantize_output=quantize_output, n_bits_A=n_bits_A)

    @classmethod
    def from_quant_args(cls, module, args):
        """empty module with the recorded constructor arguments (weights come from the checkpoint's state dict)"""
        return cls(module.in_features, module.out_features, module.bias is not None, device=module.weight.device, **args)

    @staticmethod
    def from_float(module, init_only=False, weight_quant='per_channel', act_quant='per_token', quantize_output=False,
                   n_bits_W=8, n_bits_A=16, group_size_W=0, codeBookQuantInd=False, debugPath=[], debug=False):
        assert isinstance(module, torch.nn.Linear)
        new_module = WxAxLinear(module.in_features, module.out_features, module.bias is not None,
                                weight_quant=weight_quant, act_quant=act_quant, quantize_output=quantize_output,
                                n_bits_A=n_bits_A, device=module.weight.device)
        if init_only:
            return new_module
        if weight_quant == 'per_channel':
            new_module.weight.data.copy_(quantize_weight_per_channel_absmax(module.weight.data, n_bits_W))
        elif weight_quant == 'per_tensor':
            new_module.weight.data.copy_(quantize_weight_per_tensor_absmax(module.weight.data, n_bits_W))
        elif weight_quant == 'group':
            new_module.weight.data.copy_(quantize_weight_absmax(module.weight.data, n_bits_W, group_size_W,
                                                                codeBookQuantInd=codeBookQuantInd,
                                                                debugPath=debugPath, debug=debug))
        else:
            raise ValueError(f'Invalid weight_quant: {weight_quant}')
        new_module.weight_quant_name = weight_quant
        if module.bias is not None:
            new_module.bias.data.copy_(module.bias.data.to(new_module.weight.dtype))
        return new_module

    def quant_args(self):
        """constructor arguments that are not in the state dict: what a packed checkpoint must carry to rebuild this module
        (models.save_quantized / from_quantized)"""
        return {"weight_quant": self.weight_quant_name, "act_quant": self.act_quant_name, "quantize_output": self.quantize_output,
                "n_bits_A": self.n_bits_A, "q_act": self.quantize_act}

    def __repr__(self):
        return (f'WxAxLinear({self.in_features}, {self.out_features}, bias={self.bias is not None}, '
                f'weight_quant={self.weight_quant_name}, act_quant={self.act_quant_name}, '
                f'output_quant={self.output_quant_name})')


class WxAxConv2d(nn.Module):
    """fake_quant.py:263-398: nn.Conv2d replacement.  The weight RTN runs in kernel (b); the convolution
    itself stays cuDNN (outside the north-star kernels a-d)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 act_group_size=1, weight_quant='per_tensor', act_quant='per_token', quantize_output=False, n_bits_A=16,
                 device=None):
        super().__init__()
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = pair(kernel_size), pair(stride)
        self.padding, self.dilation = pair(padding), pair(dilation)
        self.groups = groups
        self.a_gs = act_group_size
        self.quantise_act = quantize_output
        self.n_bits_A = n_bits_A
        assert in_channels % groups == 0
        self.register_buffer('weight', torch.empty((out_channels, in_channels // groups, *self.kernel_size),
                                                   dtype=torch.float16, device=device))
        if bias:
            self.register_buffer('bias', torch.zeros(out_channels, dtype=torch.float16, device=device))
        else:
            self.register_buffer('bias', None)
        self.weight_quant_name = weight_quant
        table = {
            'per_token': partial(quantize_activation_per_token_absmax, n_bits=n_bits_A),
            'per_tensor': partial(quantize_activation_per_tensor_absmax, n_bits=n_bits_A),
            'per_channel': partial(quantize_activation_per_channel_absmax, n_bits=n_bits_A),
            'per_group': partial(quantize_activation_per_channel_group_absmax, n_bits=n_bits_A, group_size=self.a_gs),
        }
        if act_quant not in table:
            raise ValueError(f'Invalid act_quant: {act_quant}')
        self.act_quant_name, self.act_quant = act_quant, table[act_quant]
        if quantize_output:
            self.output_quant_name, self.output_quant = self.act_quant_name, self.act_quant
        else:
            self.output_quant_name, self.output_quant = 'None', (lambda x: x)

    @torch.no_grad()
    def forward(self, x):
        q_x = self.act_quant(x) if self.quantise_act else x
        w = self.weight if self.weight.dtype == q_x.dtype else self.weight.to(q_x.dtype)
        b = self.bias if (self.bias is None or self.bias.dtype == q_x.dtype) else self.bias.to(q_x.dtype)
        if self._pointwise_gemm(q_x):
            # 1x1 convolution = F.linear over channels: the tcgen05 GEMM on the [B*H*W, C] token view
            # (SURVEY.md section 8(f) row 3); the result is the channels-last tensor cuDNN would return.
            from .linear import nchw_as_tokens, tokens_as_nchw
            n, _, h, wd = q_x.shape
            y = tokens_as_nchw(ops.gemm_f16(nchw_as_tokens(q_x), w.reshape(self.out_channels, self.in_channels), b), n, h, wd)
        elif self.conv3x3_gemm and self._conv3x3_gemm(q_x):
            # 3x3 / pad 1: one implicit GEMM (qdm_conv3x3_f16 / _nhwc_f16 at stride 1, qdm_conv3x3s2_nhwc_f16 at stride 2)
            y = ops.conv3x3_f16(q_x, self._taps(w), b, stride=self.stride[0])
        else:
            y = torch.nn.functional.conv2d(q_x, w, b, self.stride, self.padding, self.dilation, self.groups)
        return self.output_quant(y).to(x.dtype)

    # implicit-GEMM 3x3 path; class-level switch so that A/B timing against cuDNN is one line
    # Off by default for the fake-quant module: with fp16 weights there is no memory to save and cuDNN's channels-last
    # kernels are 15-25 % faster than the implicit GEMM (profiles/conv3x3_r01.json); QDM_CONV_GEMM=1 or setting the
    # class attribute turns it on.  The packed-int4 module (linear.QConv3x3) always runs the implicit GEMM.
    conv3x3_gemm = os.environ.get("QDM_CONV_GEMM", "0") == "1"

    def _conv3x3_gemm(self, x):
        if self.stride == (2, 2) and not (x.dim() == 4 and ops.conv3x3_stride2_ok(x.shape[2], x.shape[3])):
            return False
        return (self.kernel_size == (3, 3) and self.stride in ((1, 1), (2, 2)) and self.padding == (1, 1)
                and self.dilation == (1, 1) and self.groups == 1 and x.dim() == 4 and x.is_cuda
                and x.dtype in (torch.float16, torch.bfloat16)
                and self.in_channels % 64 == 0 and self.out_channels % 8 == 0)

    def _taps(self, w):
        """[N, 9C] tap-major copy of the (fake-quant) weight, rebuilt only when the weight buffer changes"""
        key = (w.data_ptr(), w._version, w.dtype)
        if getattr(self, "_taps_key", None) != key:
            self._taps_cache, self._taps_key = ops.conv3x3_weight_taps(w), key
        return self._taps_cache

    def _pointwise_gemm(self, x):
        return (self.kernel_size == (1, 1) and self.stride == (1, 1) and self.padding == (0, 0)
                and self.dilation == (1, 1) and self.groups == 1 and x.dim() == 4 and x.is_cuda
                and x.dtype in (torch.float16, torch.bfloat16)
                and self.in_channels % 8 == 0 and self.out_channels % 8 == 0)

    @classmethod
    def from_float(cls, module, init_only=False, weight_quant='per_tensor', act_quant='per_tensor', act_group_size=1,
                   quantize_output=False, n_bits_W=8, n_bits_A=16, group_size_W=0, codeBookQuantInd=False,
                   debugPath=[], debug=False):
        assert isinstance(module, torch.nn.Conv2d)
        new_module = cls(module.in_channels, module.out_channels, module.kernel_size, module.stride, module.padding,
                         module.dilation, module.groups, module.bias is not None, act_quant=act_quant,
                         quantize_output=quantize_output, n_bits_A=n_bits_A, act_group_size=act_group_size,
                         device=module.weight.device)
        if init_only:
            return new_module
        if weight_quant == 'per_channel':
            new_module.weight.data.copy_(quantize_weight_per_channel_absmax(module.weight.data, n_bits_W))
        elif weight_quant == 'per_tensor':
            new_module.weight.data.copy_(quantize_weight_per_tensor_absmax(module.weight.data, n_bits_W))
        elif weight_quant == 'group':
            new_module.weight.data.copy_(quantize_weight_absmax(module.weight.data, n_bits_W, group_size_W,
                                                                codeBookQuantInd=codeBookQuantInd,
                                                                debugPath=debugPath, debug=debug))
        else:
            raise ValueError(f'Invalid weight_quant: {weight_quant}')
        new_module.weight_quant_name = weight_quant
        if module.bias is not None:
            new_module.bias.data.copy_(module.bias.data.to(new_module.weight.dtype))
        return new_module

    def quant_args(self):
        """constructor arguments that are not in the state dict (see WxAxLinear.quant_args)"""
        return {"weight_quant": self.weight_quant_name, "act_quant": self.act_quant_name, "act_group_size": self.a_gs,
                "quantize_output": self.quantise_act, "n_bits_A": self.n_bits_A}

    @classmethod
    def from_quant_args(cls, module, args):
        return cls(module.in_channels, module.out_channels, module.kernel_size, module.stride, module.padding, module.dilation,
                   module.groups, module.bias is not None, device=module.weight.device, **args)

    def __repr__(self):
        s = f'WxAxConv2d({self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}'
        if self.padding != (0, 0):
            s += f', padding={self.padding}'
        if self.dilation != (1, 1):
            s += f', dilation={self.dilation}'
        if self.groups != 1:
            s += f', groups={self.groups}'
        if self.bias is None:
            s += ', bias=False'
        return s + (f', weight_quant={self.weight_quant_name}, act_quant={self.act_quant_name}, '
                    f'output_quant={self.output_quant_name})')
