"""Mirror of the reference's quantize/scale.py: fold AWQ scales into (prev_op, layers) pairs and clamp
weights for clipping.  K-length vector algebra on the device the module already lives on (no kernel is
needed for these) -- and none of the reference's per-call GPU<->CPU shuttling
(scale.py:42-46,81-84).
"""
from typing import List, Tuple

import torch
import torch.nn as nn

from .module import get_op_by_name

allowed_norms = [nn.LayerNorm]


class AdaLNShift:
    """Marker for norms whose affine comes from a modulation Linear (SD3.5 AdaLayerNormZero): the scale is
    folded into the rows of that Linear with the `+1, /s, -1` trick the reference uses for Gemma's
    (1 + weight) norm (scale.py:96-99).  `rows_scale` / `rows_shift` are slices of the Linear's output."""

    def __init__(self, linear: nn.Linear, rows_shift: slice, rows_scale: slice):
        self.linear, self.rows_shift, self.rows_scale = linear, rows_shift, rows_scale


def describe_prev_op(module, prev_op):
    """name of a prev_op relative to `module`; AdaLN targets become a picklable tuple so that search results
    can be gathered across ranks: ("adaln", linear_name, (shift rows), (scale rows))."""
    from .module import get_op_name
    if isinstance(prev_op, AdaLNShift):
        return ("adaln", get_op_name(module, prev_op.linear), (prev_op.rows_shift.start, prev_op.rows_shift.stop),
                (prev_op.rows_scale.start, prev_op.rows_scale.stop))
    return get_op_name(module, prev_op) if isinstance(prev_op, nn.Module) else prev_op


def resolve_prev_op(module, desc):
    if isinstance(desc, str):
        return get_op_by_name(module, desc)
    if isinstance(desc, (tuple, list)) and len(desc) == 4 and desc[0] == "adaln":
        return AdaLNShift(get_op_by_name(module, desc[1]), slice(*desc[2]), slice(*desc[3]))
    return desc


@torch.no_grad()
def apply_clip(module, clip_list: Tuple[str, torch.Tensor]):
    """scale.py:25-34."""
    for name, max_val in clip_list:
        layer: nn.Linear = get_op_by_name(module, name)
        max_val = max_val.to(layer.weight.device)
        org_shape = layer.weight.shape
        w = layer.weight.data.reshape(*max_val.shape[:2], -1)
        layer.weight.data = torch.clamp(w, -max_val, max_val).reshape(org_shape)


def apply_scale(module, scales_list, input_feat_dict=None):
    """scale.py:37-84."""
    for prev_op_name, layer_names, scales in scales_list:
        prev_op = resolve_prev_op(module, prev_op_name)
        layers = [get_op_by_name(module, name) for name in layer_names]
        scales = scales.to(layers[0].weight.device)
        if isinstance(prev_op, AdaLNShift):
            scale_adaln_fcs(prev_op, layers, scales)
        elif isinstance(prev_op, nn.Linear) and isinstance(layers, list) and isinstance(layers[0], nn.Linear):
            scale_fc_fcs(prev_op, layers, scales)
        elif any(isinstance(prev_op, t) for t in allowed_norms) or "rmsnorm" in str(prev_op.__class__).lower():
            scale_ln_fcs(prev_op, layers, scales)
        else:
            raise NotImplementedError(f"prev_op {type(prev_op)} not supported yet!")
        if input_feat_dict is not None:
            for layer_name in layer_names:
                if layer_name in input_feat_dict:
                    inp = input_feat_dict[layer_name]
                    inp.div_(scales.view(1, -1).to(inp.device))


@torch.no_grad()
def scale_ln_fcs(ln, fcs: List[nn.Linear], scales: torch.Tensor):
    """scale.py:88-114 (plain-weight norms)."""
    if not isinstance(fcs, list):
        fcs = [fcs]
    scales = scales.to(ln.weight.device)
    ln.weight.div_(scales)
    if hasattr(ln, "bias") and ln.bias is not None:
        ln.bias.div_(scales)
    for fc in fcs:
        fc.weight.mul_(scales.view(1, -1))
    for p in list(ln.parameters()) + [q for fc in fcs for q in fc.parameters()]:
        assert torch.isnan(p).sum() == 0


@torch.no_grad()
def scale_adaln_fcs(ada: AdaLNShift, fcs: List[nn.Linear], scales: torch.Tensor):
    """x_mod = LN(x) * (1 + scale) + shift, with [shift | scale] = ada.linear(emb).  Dividing x_mod by s means
    shift/s and (1 + scale)/s - 1 = scale/s + (1/s - 1): rows of the modulation Linear are divided by s and the
    scale rows' bias gets the (1/s - 1) offset -- the same algebra as the Gemma branch of scale.py:96-99."""
    lin = ada.linear
    s = scales.to(lin.weight.device).to(lin.weight.dtype)
    lin.weight[ada.rows_shift].div_(s.view(-1, 1))
    lin.weight[ada.rows_scale].div_(s.view(-1, 1))
    if lin.bias is not None:
        lin.bias[ada.rows_shift].div_(s)
        b = lin.bias[ada.rows_scale]
        b += 1
        b.div_(s)
        b -= 1
    for fc in fcs:
        fc.weight.mul_(s.view(1, -1))


@torch.no_grad()
def scale_fc_fc(fc1: nn.Linear, fc2: nn.Linear, scales: torch.Tensor):
    """scale.py:117-133."""
    scale_fc_fcs(fc1, [fc2], scales)


@torch.no_grad()
def scale_fc_fcs(fc1: nn.Linear, fcs: List[nn.Linear], scales: torch.Tensor):
    """scale.py:136-153."""
    if not isinstance(fcs, list):
        fcs = [fcs]
    scales = scales.to(fc1.weight.device)
    fc1.weight[-scales.size(0):].div_(scales.view(-1, 1))
    if fc1.bias is not None:
        fc1.bias[-scales.size(0):].div_(scales.view(-1))
    for fc in fcs:
        fc.weight.mul_(scales.view(1, -1))
    for p in list(fc1.parameters()) + [q for fc in fcs for q in fc.parameters()]:
        assert torch.isnan(p).sum() == 0
