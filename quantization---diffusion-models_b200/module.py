"""Module-tree helpers with the semantics of the reference's utils/module.py (names kept so the quantizer
code reads the same)."""
import torch
import torch.nn as nn


def get_named_linears(module):
    """utils/module.py:12."""
    return {n: m for n, m in module.named_modules() if isinstance(m, nn.Linear)}


def get_named_linears_and_conv_layers(module):
    """utils/module.py:15."""
    return {n: m for n, m in module.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))}


def get_op_by_name(module, op_name):
    """utils/module.py:18-23."""
    found = dict(module.named_modules()).get(op_name)
    if found is None:
        raise ValueError(f"Cannot find op {op_name} in module {module}")
    return found


def set_op_by_name(layer, name, new_module):
    """utils/module.py:26-37: dotted path, digits index into containers."""
    *path, leaf = name.split(".")
    for part in path:
        layer = layer[int(part)] if part.isdigit() else getattr(layer, part)
    setattr(layer, leaf, new_module)


def get_op_name(module, op):
    """utils/module.py:40-45."""
    for n, m in module.named_modules():
        if m is op:
            return n
    raise ValueError(f"Cannot find op {op} in module {module}")


def append_str_prefix(x, prefix):
    """utils/module.py:48-56."""
    if isinstance(x, str):
        return prefix + x
    if isinstance(x, (tuple, list)):
        return type(x)(append_str_prefix(y, prefix) for y in x)
    return x


def exclude_layers_to_not_quantize(linear_layers, modules_to_not_convert):
    """utils/module.py:59-67."""
    if modules_to_not_convert is None:
        return linear_layers
    return {n: l for n, l in linear_layers.items() if not any(key in n for key in modules_to_not_convert)}


class ModuleTraversal:
    """utils/module.py:69-86 (== AwqQuantizer.MyTraversal, quantize/quantizer.py:142-159): depth-first list of
    (parent, attribute name, layer) for every nn.Linear / nn.Conv2d."""

    def __init__(self):
        self.name, self.parent, self.lin_conv = None, None, []

    def traverse(self, name, module, parent):
        self.name, self.parent = name, parent
        if isinstance(module, (torch.nn.Linear, torch.nn.Conv2d)):
            self.lin_conv.append((parent, name, module))
        for child_name, child in module.named_children():
            if child is not None:
                self.traverse(child_name, child, module)

    def get_lin_conv(self):
        return self.lin_conv


def get_lin_conv_layers(name, module, root):
    """utils/module.py:88-92."""
    t = ModuleTraversal()
    t.traverse(name, module, root)
    return t.get_lin_conv()
