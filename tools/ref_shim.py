"""Import the UNMODIFIED reference (/root/reference) in this container.

The reference is an overlay on AutoAWQ: every intra-package import is `awq.*` and several
third-party imports are absent here.  This shim registers an `awq` namespace whose sub-packages
point at the reference directories and stubs the absent modules, so that
`awq.quantize.quantizer`, `awq.quantize.quantizer_SQ`, `awq.quantize.fake_quant`,
`awq.quantize.scale`, `awq.utils.*` import as they lie.  Nothing is copied.

Used ONLY by tools/gen_golden.py and tests that compare the oracle with the live reference
(skipped when /root/reference is absent, i.e. on the GPU box).
"""
import importlib
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("QDM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "quantize"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _namespace(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    m.__spec__.submodule_search_locations = [path]
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Idempotent. Returns True when the reference can be imported."""
    global _installed
    if _installed:
        return True
    if not available():
        return False
    import torch

    _namespace("awq", REF_ROOT)
    for sub in ("quantize", "utils", "models", "modules"):
        _namespace(f"awq.{sub}", os.path.join(REF_ROOT, sub))

    class _Dummy(torch.nn.Module):
        pass

    def _try(name):
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not _try("hadamard_transform"):
        _stub("hadamard_transform", hadamard_transform=lambda x: x)
    if not _try("fast_pytorch_kmeans"):
        _stub("fast_pytorch_kmeans", KMeans=object)
    if not _try("matplotlib"):
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    if not _try("torchsummary"):
        _stub("torchsummary", summary=lambda *a, **k: None)
    if not _try("datasets"):
        _stub("datasets", load_dataset=lambda *a, **k: None)
    _try("transformers")  # must see the real environment before `accelerate` is stubbed
    for sub in ("bloom.modeling_bloom", "llama.modeling_llama", "gemma.modeling_gemma", "gemma2.modeling_gemma2",
                "cohere.modeling_cohere"):
        _try(f"transformers.models.{sub}")
    _try("transformers.activations")
    if not _try("accelerate"):
        _stub("accelerate")
    if not _try("diffusers"):
        d = _stub("diffusers", DiffusionPipeline=object)
        cb = _stub("diffusers.callbacks", PipelineCallback=object)
        d.callbacks = cb
    _stub("awq.modules.linear", WQLinear_GEMM=type("WQLinear_GEMM", (_Dummy,), {}),
          WQLinear_GEMV=type("WQLinear_GEMV", (_Dummy,), {}),
          WQLinear_Marlin=type("WQLinear_Marlin", (_Dummy,), {}),
          WQLinear_GEMVFast=type("WQLinear_GEMVFast", (_Dummy,), {}))
    _stub("awq.modules.act", ScaledActivation=type("ScaledActivation", (_Dummy,), {}))
    _stub("awq.models.base", diffusers=sys.modules.get("diffusers"))
    _installed = True
    return True


def ref():
    """Namespace object with the reference modules the hot path lives in."""
    if not install():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    ns = types.SimpleNamespace()
    ns.fake_quant = importlib.import_module("awq.quantize.fake_quant")
    ns.quantizer = importlib.import_module("awq.quantize.quantizer")
    ns.quantizer_SQ = importlib.import_module("awq.quantize.quantizer_SQ")
    ns.scale = importlib.import_module("awq.quantize.scale")
    ns.packing_utils = importlib.import_module("awq.utils.packing_utils")
    ns.quant_utils = importlib.import_module("awq.utils.quant_utils")
    ns.calib_data = importlib.import_module("awq.utils.calib_data")
    ns.module = importlib.import_module("awq.utils.module")
    return ns
