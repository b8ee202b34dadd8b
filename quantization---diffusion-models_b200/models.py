"""Top-level API with the reference's signatures for the quantized-linear path:
`BaseAWQForDiffusion.quantize(tokenizer=None, quant_config={...}, quantType='awq'|'sq', quantUnet=..., ...)`
(models/base.py:216-526) and `.generate(prompt, ..., lat=..., generator=...)` (models/base.py:829-850), plus the
three adapters `StableDiffusion1_x`, `StableDiffusionXL`, `StableDiffusion3_5`
(models/StableDiffusion1_x.py, StableDiffusionXL.py, StableDiffusion3_5.py).

`from_pretrained` needs diffusers + a checkpoint, neither of which exists offline; `from_skeleton` builds the
same module tree with random weights (skeletons.py).  `save_quantized` / `from_quantized` store the REAL packed
tensors (qweight/qzeros/scales) -- SURVEY.md section 8(f) row 2 -- instead of fp16 fake-quant weights.
"""
import json
import os
from typing import Dict

import torch
import torch.nn as nn

from .calib_data import Input_Capture_Hook
from .config import AwqConfig
from .quantizer import AwqQuantizer
from .quantizer_SQ import SqQuantizer
from .scale import AdaLNShift
from . import skeletons as sk


class BaseAWQForDiffusion:
    kind = None

    def __init__(self, pipeline, model_type, is_quantized=False, config=None, quant_config=None):
        """models/base.py:120-138."""
        self.pipeline = pipeline
        self.model_type, self.is_quantized, self.config = model_type, is_quantized, config or {}
        self.quant_config: AwqConfig = quant_config or AwqConfig()
        self.search_result = None
        self.quantized_components = []
        self.quantizer = None
        self.calib_samples = None          # list of (prompts, latents); defaults to a small synthetic set
        self.calib_steps = 4               # denoise steps per calibration batch for the AWQ capture pass
        self.calib_max_tokens = 4096       # tokens kept per Linear input (subsampled evenly)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, model_path, model_type=None, torch_dtype=torch.float16, **kwargs):
        raise RuntimeError("from_pretrained needs `diffusers` and a checkpoint; neither is available offline. "
                           "Use from_skeleton() (random-init weights of the same architecture).")

    @classmethod
    def from_skeleton(cls, device="cuda", dtype=torch.float16, seed=42, **arch):
        """Random-init weights of the architecture.  On a CUDA device the modules are created there directly in `dtype`
        (the full SD3.5-L skeleton is 8 G parameters: initialising it on the host in fp32 takes minutes and 32 GB); the
        CUDA generator is seeded, so every rank of a multi-GPU run builds the identical model."""
        torch.manual_seed(seed)
        dev = torch.device(device)
        if dev.type == "cuda":
            old = torch.get_default_dtype()
            torch.set_default_dtype(dtype)
            try:
                with torch.device(dev):
                    den = cls.build_denoiser(**arch)
            finally:
                torch.set_default_dtype(old)
        else:
            den = cls.build_denoiser(**arch)
        pipe = sk.SkeletonPipeline(cls.kind, den, device=device, dtype=dtype, latent_size=arch.get("latent_size"))
        return cls(pipe, cls.kind, False, {"arch": arch}, AwqConfig())

    def to(self, device):
        return self.pipeline.to(device)

    # ------------------------------------------------------------------ adapter surface (StableDiffusion1_x.py:39-102)
    def denoiser(self):
        return self.pipeline.unet if self.pipeline.unet is not None else self.pipeline.transformer

    def get_model_layers_unet(self):
        if self.pipeline.unet is None:
            raise Exception("There is no UNet in this model")
        return [list(self.pipeline.unet.named_children())]

    def get_model_layers_transformers(self):
        if self.pipeline.transformer is None:
            raise Exception("There is no transformer in this model")
        return [list(self.pipeline.transformer.named_children())]

    def get_model_layers_te(self):
        return []

    def get_model_layers_vae(self):
        return []

    def get_root(self, component, idx):
        return self.denoiser()

    def get_unet(self):
        return self.pipeline.unet

    def get_transformer(self):
        return self.pipeline.transformer

    def get_pipeline(self):
        return self.pipeline

    def set_quantized_components(self, component):
        self.quantized_components.append(component)

    def get_debugModuleNames(self, **_):
        return []

    def get_scalingStates(self, **_):
        return []

    def get_projectionNames(self, **_):
        return []

    def mean_of_dict(self, act_dict):
        """models/StableDiffusion1_x.py:104-112: mean over calls of the per-call maxima.  The fused hook
        (calib_data.Fused_Mean_Max_Activation_Hook) hands over the accumulated sum instead of a dict of tensors."""
        if hasattr(act_dict, "mean_over_calls"):
            return act_dict.mean_over_calls()
        return torch.mean(torch.stack(list(act_dict.values())), dim=0)

    # ------------------------------------------------------------------ AWQ search support (new for diffusion)
    def block_types(self):
        return (sk.BasicTransformerBlock,)

    def get_search_blocks(self):
        return {n: m for n, m in self.denoiser().named_modules() if isinstance(m, self.block_types())}

    get_smoothing_blocks = get_search_blocks   # models/StableDiffusion1_x.py:96-102

    @staticmethod
    def block_cost(block):
        return sum(m.weight.numel() for m in block.modules() if isinstance(m, nn.Linear))

    def default_calib_samples(self, n_batches=1, batch_size=2):
        from .calib_data import get_calib_dataset_dm
        return get_calib_dataset_dm(self.pipeline, "synthetic-captions", batch_size=batch_size,
                                    n_samples=n_batches * batch_size, seed=42, device=self.pipeline.device)

    @torch.no_grad()
    def capture_block_inputs(self, block_names, shard=None, wanted_by=None):
        """FP pass over the calibration set with a capture hook on every Linear of the requested blocks:
        {block: {linear_name: X [n_tok, K]}} kept on the GPU (the FP activations of the un-quantised model, as
        quantizer.py:1093-1141 collects them block by block), at most `calib_max_tokens` rows per Linear.
        shard = (rank, world): the pass is DATA PARALLEL -- calibration batch i runs on rank i % world with hooks on every
        block some rank searches (`wanted_by`: {block: [ranks]}), then dist.exchange_captures hands each rank the inputs
        of `block_names` from all batches in single-process order.  A batch is always run whole by one rank, so every
        forward has the shapes (and cuBLAS kernels) of the single-process run: captured inputs, and with them the
        searched scales and packed codes, are identical for every world size."""
        import time
        from .dist import exchange_captures
        t_in = time.perf_counter()
        rank, world = shard if shard is not None else (0, 1)
        blocks = self.get_search_blocks()
        samples = self.calib_samples or self.default_calib_samples()
        per_call = -(-self.calib_max_tokens // max(1, len(samples) * self.calib_steps))   # ceil(max_tokens / total calls)
        hook_blocks = list(wanted_by) if (world > 1 and wanted_by) else list(block_names)
        hooks = {}
        n_local = sum(1 for bi in range(len(samples)) if bi % world == rank) * self.calib_steps   # forward calls this rank runs
        lins = [(bn, ln, lin) for bn in hook_blocks for ln, lin in blocks[bn].named_modules() if isinstance(lin, nn.Linear)]
        # ONE arena for every kept row of this pass (upper bound per_call rows per call and Linear), sliced per hook
        par = next(self.denoiser().parameters())
        rows = n_local * per_call
        arena = torch.empty(sum(rows * lin.in_features for _, _, lin in lins), dtype=par.dtype, device=par.device) if lins else None
        off = 0
        for bn, ln, lin in lins:
            h = Input_Capture_Hook(self.calib_max_tokens, per_call, arena=arena[off:off + rows * lin.in_features].view(rows, lin.in_features))
            off += rows * lin.in_features
            h.hook_handle = lin.register_forward_hook(h)
            hooks[(bn, ln)] = h
        on_gpu = torch.device(self.pipeline.device).type == "cuda"
        sync = torch.cuda.synchronize if on_gpu else (lambda: None)
        t0 = time.perf_counter()
        for bi, (prompts, latents) in enumerate(samples):
            if bi % world != rank:
                continue
            for h in hooks.values():
                h.next_id = bi * self.calib_steps          # global index of this batch's first forward call
            self.pipeline(prompt=prompts, latents=latents, num_inference_steps=self.calib_steps, guidance_scale=7.5)
        caps = {bn: {} for bn in hook_blocks}
        for (bn, ln), h in hooks.items():
            h.hook_handle.remove()
            caps[bn][ln] = h.chunks
        sync()
        t1 = time.perf_counter()
        if world > 1:
            caps = exchange_captures(caps, wanted_by or {bn: [rank] for bn in block_names}, rank, world, self.pipeline.device)
            sync()
        t2 = time.perf_counter()
        out = {bn: {ln: Input_Capture_Hook.merge(ch, self.calib_max_tokens) for ln, ch in caps[bn].items()} for bn in block_names}
        sync()
        self.capture_timings = {"capture_hooks_s": t0 - t_in, "capture_forward_s": t1 - t0, "capture_exchange_s": t2 - t1,
                                "capture_merge_s": time.perf_counter() - t2}
        return out

    def get_layers_for_scaling(self, block, input_feat):
        """Scaling groups of a BasicTransformerBlock: the two groups the reference's SmoothQuant
        table has (StableDiffusion1_x.py:117-139) plus norm2 -> attn2.to_q.  Each group: prev_op, layers, inp,
        module2inspect (the dict shape of the LLM adapters, e.g. models/llava.py:42-89)."""
        class _Cat(nn.Module):
            def __init__(self, ls):
                super().__init__()
                self.ls = nn.ModuleList(ls)

            def forward(self, x):
                return torch.cat([l(x) for l in self.ls], dim=-1)

        qkv = [block.attn1.to_q, block.attn1.to_k, block.attn1.to_v]
        return [
            dict(prev_op=block.norm1, layers=qkv, inp=input_feat["attn1.to_q"], module2inspect=_Cat(qkv)),
            dict(prev_op=block.norm2, layers=[block.attn2.to_q], inp=input_feat["attn2.to_q"]),
            dict(prev_op=block.norm3, layers=[block.ff.net[0].proj], inp=input_feat["ff.net.0.proj"]),
        ]

    def get_layers_for_scaling_unet(self, module, hooks):
        """models/StableDiffusion1_x.py:115-150 (SmoothQuant groups with their activation statistic)."""
        return [
            dict(prev_op=module.norm1, layers=[module.attn1.to_q, module.attn1.to_k, module.attn1.to_v],
                 activations_max=[self.mean_of_dict(hooks[n].max_scales) for n in ('attn1.to_q', 'attn1.to_k', 'attn1.to_v')]),
            dict(prev_op=module.norm3, layers=[module.ff.net[0].proj],
                 activations_max=[self.mean_of_dict(hooks['ff.net.0.proj'].max_scales)]),
        ]

    # ------------------------------------------------------------------ models/base.py:216-526
    @torch.no_grad()
    def quantize(self, tokenizer=None, quant_config: Dict = {}, calib_data="pileval", split="train", text_column="text",
                 duo_scaling=True, export_compatible=False, apply_clip=True, applyScale=True, quant_act=False,
                 n_parallel_calib_samples=None, max_chunk_memory=1024 * 1024 * 1024, quantType="awq", quantUnet=True,
                 quantTextEncoder=False, quantVAE=False, quantTransformer=False, codeBookQuantInd=False,
                 debugSavePath="", debugPlot=False, calibrate=False, alpha=0.5, shard=None, **kwargs):
        """Same call shape as the reference (first positional is `tokenizer`; the method is picked by `quantType`).
        New keywords: `calibrate` (run the AWQ scale/clip search on the diffusion blocks), `alpha` (SmoothQuant),
        `shard` = (rank, world) for the sharded search."""
        if hasattr(self.pipeline, "invalidate_graphs"):
            self.pipeline.invalidate_graphs()          # captured denoise steps point at the modules about to be replaced
        self.unfuse_layers()                           # fused copies of the packed tensors about to be replaced
        quant_config = dict(quant_config)
        if quant_act and quant_config.get('version', 'fake_act').lower() != 'fake_act':
            quant_config['version'] = 'fake_act'
        self.quant_config = AwqConfig.from_dict(quant_config)
        qc = self.quant_config
        common = dict(quantise_act=qc.quantize_act, weight_quant_conv_type=qc.weight_quant_conv_type,
                      weight_quant_type=qc.weight_quant_type, act_quant_conv_type=qc.act_quant_conv_type,
                      act_quant_conv_group_size=qc.act_quant_conv_group_size, w_bit=qc.w_bit, wv_bit=qc.wv_bit, a_bit=qc.a_bit,
                      group_size=qc.q_group_size, zero_point=qc.zero_point, version=qc.version, calib_data=calib_data,
                      split=split, text_column=text_column, duo_scaling=duo_scaling,
                      modules_to_not_convert=qc.modules_to_not_convert, export_compatible=export_compatible,
                      quant_act=quant_act, apply_clip=apply_clip, applyScale=applyScale,
                      n_parallel_calib_samples=n_parallel_calib_samples, max_chunk_memory=max_chunk_memory,
                      quantUnet=quantUnet and self.pipeline.unet is not None, quantTextEncoder=quantTextEncoder, quantVAE=quantVAE,
                      quantTransformer=quantTransformer or (self.pipeline.transformer is not None and quantUnet),
                      diffusion_model=True, codeBookQuantInd=codeBookQuantInd)
        if quantType.lower() == 'awq':
            self.quantizer = AwqQuantizer(self, None, None, calibrate=calibrate, **common)
            if shard is not None and calibrate:
                from .dist import sharded_search
                results = sharded_search(self.quantizer, shard)
                self.quantizer.apply_search_results(results)
                self.quantizer.calibrate = False
            self.quantizer.quantize(debugSavePath, debugPlot)
        elif quantType.lower() == 'sq':
            self.quantizer = SqQuantizer(self, None, None, alpha=alpha, **{**common, **kwargs})
            self.quantizer.quantize(debugSavePath, debugPlot, samples=self.calib_samples, shard=shard)
        else:
            raise NotImplementedError("Only awq and sq are supported for now.")
        self.is_quantized = True

    # ------------------------------------------------------------------ models/base.py:829-850
    @torch.no_grad()
    def fuse_layers(self):
        """Same-input packed projections as one launch each (fused_utils.fuse_projections; the reference's `fuse_layers`
        switch of from_quantized, models/base.py:736-826, names the same transformation for its LLM blocks).
        Returns {kind: members fused}; a no-op for models without packed Linears."""
        from .fused_utils import fuse_projections
        if hasattr(self.pipeline, "invalidate_graphs"):
            self.pipeline.invalidate_graphs()
        self._fused = fuse_projections(self.denoiser())
        return self._fused

    def unfuse_layers(self):
        from .fused_utils import unfuse_projections
        if self.pipeline is not None and getattr(self, "_fused", None):
            unfuse_projections(self.denoiser())
            if hasattr(self.pipeline, "invalidate_graphs"):
                self.pipeline.invalidate_graphs()
        self._fused = None

    def generate(self, prompt, height=512, width=512, num_inference_steps=50, guidance_scale=7.5, negative_prompt=None,
                 num_images_per_prompt=1, generator=None, device="cpu", lat=None, output_type=None, cuda_graph=False,
                 fuse_layers=False, **kwargs):
        """models/base.py:829-850.  `cuda_graph=True` replays each denoiser call from a CUDA graph (skeletons.SkeletonPipeline);
        `fuse_layers=True` runs the same-input packed projections as one launch each (fuse_layers(), built once)."""
        if self.pipeline is None:
            raise RuntimeError("The diffusion pipeline is not loaded. Please use `from_pretrained` or `from_quantized` first.")
        if fuse_layers and not getattr(self, "_fused", None):
            self.fuse_layers()
        return self.pipeline(prompt=prompt, num_inference_steps=num_inference_steps, guidance_scale=guidance_scale,
                             num_images_per_prompt=1, generator=generator, latents=lat, output_type=output_type, cuda_graph=cuda_graph)

    # ------------------------------------------------------------------ packed checkpoint
    def save_quantized(self, save_dir):
        """models/base.py:530-582 re-thought: the denoiser state dict (packed int4 / int8 tensors for swapped
        modules) + quantization_config + the list of quantised components."""
        os.makedirs(save_dir, exist_ok=True)
        kinds, wrapped = {}, set()
        for n, m in self.denoiser().named_modules():
            k = type(m).__name__
            if k == "QConv1x1":                      # pointwise conv on the GEMM kernels: record the inner module kind
                kinds[n] = {"kind": f"QConv1x1:{type(m.inner).__name__}",
                            "args": {"w_bit": getattr(m.inner, "w_bit", 8), "group_size": getattr(m.inner, "group_size", 0)}}
                wrapped.add(n + ".inner")
            elif k in ("WQLinear_GEMM", "QConv3x3") and n not in wrapped:
                kinds[n] = {"kind": k, "args": {"w_bit": m.w_bit, "group_size": m.group_size}}
            elif k == "W8A8Linear" and n not in wrapped:
                kinds[n] = {"kind": k, "args": {}}
            elif k in ("WxAxLinear", "WxAxConv2d"):
                # everything the constructor needs that the state dict does not hold: activation / output quantisers,
                # their bit width and group size (a model quantised with quantize_act=True must reload with it ON)
                kinds[n] = {"kind": k, "args": m.quant_args()}
        meta = {"model_type": self.model_type, "arch": self.config.get("arch", {}), "quantization_config": self.quant_config.to_transformers_dict(),
                "quant_config": self.quant_config.to_dict(), "quant_components": self.quantized_components, "modules": kinds}
        with open(os.path.join(save_dir, "quant_components.json"), "w") as f:
            json.dump(meta, f, indent=1)
        # the weights as safetensors (models/base.py:566-582 writes `model.safetensors` the same way): packed int32 / int8
        # tensors keep their dtype; the Hugging Face `quantization_config` (models/_config.py:97-107) also travels in the
        # file's metadata, so the file alone says how to read it.  Tensors that share storage are written once per name.
        from safetensors.torch import save_file
        state, seen = {}, set()
        for k, v in self.denoiser().state_dict().items():
            v = v.detach().contiguous()
            state[k] = v.clone() if (v.numel() and v.data_ptr() in seen) else v
            seen.add(v.data_ptr())
        save_file(state, os.path.join(save_dir, "model.safetensors"),
                  metadata={"format": "pt", "quantization_config": json.dumps(meta["quantization_config"]), "model_type": str(self.model_type)})

    @classmethod
    def from_quantized(cls, save_dir, device="cuda", dtype=torch.float16):
        """models/base.py:736-826: rebuild the module tree, swap in `init_only` quantised modules, load the state."""
        from .fake_quant import WxAxConv2d, WxAxLinear
        from .linear import QConv1x1, QConv3x3, W8A8Linear, WQLinear_GEMM, conv_group
        from .module import get_op_by_name, set_op_by_name
        from .fake_quant import _effective_group
        with open(os.path.join(save_dir, "quant_components.json")) as f:
            meta = json.load(f)
        model = cls.from_skeleton(device=device, dtype=dtype, **meta["arch"])
        model.quant_config = AwqConfig.from_dict(meta["quant_config"])
        den = model.denoiser()
        qg = model.quant_config.q_group_size
        for name, entry in meta["modules"].items():
            # entry: {"kind", "args"}; checkpoints written before the per-module arguments were recorded hold the kind only
            kind, args = (entry["kind"], entry.get("args", {})) if isinstance(entry, dict) else (entry, None)
            old = get_op_by_name(den, name)
            if kind.startswith("QConv1x1:"):
                ci, co, dev = old.in_channels, old.out_channels, old.weight.device
                if kind.endswith("WQLinear_GEMM"):
                    g = args["group_size"] if args else _effective_group(ci, qg)
                    inner = WQLinear_GEMM(args["w_bit"] if args else 4, g, ci, co, old.bias is not None, dev, dtype)
                else:
                    inner = W8A8Linear(ci, co, old.bias is not None, dev, dtype)
                new = QConv1x1(inner, ci, co)
            elif kind == "QConv3x3":
                g = args["group_size"] if args else conv_group(9 * old.in_channels, qg)
                new = QConv3x3.from_conv(old, args["w_bit"] if args else 4, g, init_only=True)
            elif kind == "WQLinear_GEMM":
                g = args["group_size"] if args else _effective_group(old.in_features, qg)
                new = WQLinear_GEMM.from_linear(old, args["w_bit"] if args else 4, g, init_only=True)
            elif kind == "W8A8Linear":
                new = W8A8Linear(old.in_features, old.out_features, old.bias is not None, old.weight.device, dtype)
            elif kind == "WxAxLinear":
                new = (WxAxLinear.from_quant_args(old, args) if args else WxAxLinear.from_float(old, init_only=True)).to(old.weight.device)
            else:
                new = (WxAxConv2d.from_quant_args(old, args) if args else WxAxConv2d.from_float(old, init_only=True)).to(old.weight.device)
            set_op_by_name(den, name, new)
        st_path = os.path.join(save_dir, "model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file
            den.load_state_dict(load_file(st_path, device=str(device)))
        else:   # checkpoints written before the safetensors format
            den.load_state_dict(torch.load(os.path.join(save_dir, "denoiser.pt"), map_location=device))
        model.is_quantized = True
        model.quantized_components = meta["quant_components"]
        return model


class StableDiffusion1_x(BaseAWQForDiffusion):
    kind = "sd15"

    @staticmethod
    def build_denoiser(channels=(320, 640, 1280, 1280), depth=(1, 1, 1, 0), ctx_dim=768, heads=8, **_):
        return sk.UNet2DConditionSkeleton(tuple(channels), tuple(depth), ctx_dim=ctx_dim, heads=heads)

    def checkQuantStatus(self, quantUnet=True, quantTextEncoder=False, quantVAE=False, quantTransformer=True):
        if quantTransformer:
            raise Exception("There is no Transformer in this Diffusion Model")


class StableDiffusionXL(BaseAWQForDiffusion):
    kind = "sdxl"

    @staticmethod
    def build_denoiser(channels=(320, 640, 1280), depth=(0, 2, 10), ctx_dim=2048, head_dim=64, add_embed_in=2816, **_):
        return sk.UNet2DConditionSkeleton(tuple(channels), tuple(depth), ctx_dim=ctx_dim, head_dim=head_dim, linear_proj=True,
                                          add_embed_in=add_embed_in)


class StableDiffusion3_5(BaseAWQForDiffusion):
    kind = "sd3"

    @staticmethod
    def build_denoiser(layers=38, dim=2432, heads=38, in_ch=16, patch=2, ctx_in=4096, pooled=2048, **_):
        return sk.MMDiTSkeleton(layers=layers, dim=dim, heads=heads, in_ch=in_ch, patch=patch, ctx_in=ctx_in, pooled=pooled)

    def block_types(self):
        return (sk.JointTransformerBlock,)

    def get_layers_for_scaling(self, block, input_feat):
        """JointTransformerBlock: the LayerNorms are non-affine and the affine comes from AdaLN modulation, so
        1/s is folded into the shift / scale rows of norm1.linear (scale.py AdaLNShift)."""
        class _Cat(nn.Module):
            def __init__(self, ls):
                super().__init__()
                self.ls = nn.ModuleList(ls)

            def forward(self, x):
                return torch.cat([l(x) for l in self.ls], dim=-1)

        d = block.attn.to_q.in_features
        qkv = [block.attn.to_q, block.attn.to_k, block.attn.to_v]
        add = [block.attn.add_q_proj, block.attn.add_k_proj, block.attn.add_v_proj]

        def rows(norm, part):
            """(shift rows, scale rows) of the modulation Linear that feed `part` ("attn" | "mlp"), by norm type:
            AdaLayerNormZero chunks (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp); AdaLayerNormContinuous
            -- norm1_context of the last, context_pre_only block -- chunks (scale, shift)."""
            if isinstance(norm, sk.AdaLayerNormContinuous):
                assert part == "attn"
                return slice(d, 2 * d), slice(0, d)
            o = 0 if part == "attn" else 3 * d
            return slice(o, o + d), slice(o + d, o + 2 * d)

        groups = [
            dict(prev_op=AdaLNShift(block.norm1.linear, *rows(block.norm1, "attn")), layers=qkv,
                 inp=input_feat["attn.to_q"], module2inspect=_Cat(qkv)),
            dict(prev_op=AdaLNShift(block.norm1_context.linear, *rows(block.norm1_context, "attn")), layers=add,
                 inp=input_feat["attn.add_q_proj"], module2inspect=_Cat(add)),
            dict(prev_op=AdaLNShift(block.norm1.linear, *rows(block.norm1, "mlp")), layers=[block.ff.net[0].proj],
                 inp=input_feat["ff.net.0.proj"]),
        ]
        if not block.context_pre_only:
            groups.append(dict(prev_op=AdaLNShift(block.norm1_context.linear, *rows(block.norm1_context, "mlp")),
                               layers=[block.ff_context.net[0].proj], inp=input_feat["ff_context.net.0.proj"]))
        return groups

    def get_layers_for_scaling_unet(self, module, hooks):
        raise NotImplementedError("SmoothQuant groups are defined for UNet BasicTransformerBlocks only (as in the reference)")
