"""Host-side mirror of the reference's quantize/quantizer.py (class AwqQuantizer): same method names,
argument meaning and results; the arithmetic runs in libqdm's sm_100a kernels.

What differs from the reference, on purpose (SURVEY.md sections 0.4, 3.2, 3.5):
  * the activation-aware scale/clip search is wired to the DIFFUSION branch (the reference hard-codes
    `calibrate = False` there, quantizer.py:1050, and only searches for LLMs): the model adapter supplies
    the transformer blocks, their captured inputs and `get_layers_for_scaling`;
  * calibration inputs stay on the GPU (the reference caches them on the CPU, quantizer.py:1099, and restores
    the weights from a CPU state dict 20 times per group, :703,743);
  * Q(W*s)/s is ONE fused kernel writing into a scratch weight; the original weight is never touched, so
    nothing has to be restored;
  * the 20 losses stay on the device and the argmin costs one sync per group (the reference syncs per chunk
    per ratio, :777);
  * `_apply_quant` honours `bitWidth` (the reference ignores it, :540-542).
(block, group) searches are independent, so `quantize()` can shard them over ranks: see dist.py.
"""
import logging
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .fake_quant import WxAxConv2d, WxAxLinear
from .linear import WQLinear_GEMM
from .module import (ModuleTraversal, exclude_layers_to_not_quantize, get_named_linears, get_op_name,
                     set_op_by_name)
from .scale import apply_clip, apply_scale


class AwqQuantizer:
    def __init__(self, awq_model, model=None, tokenizer=None, quantise_act=False, weight_quant_conv_type="per_channel",
                 weight_quant_type="group", act_quant_conv_type="per_channel", act_quant_conv_group_size=1,
                 w_bit=4, wv_bit=4, a_bit=16, group_size=128, zero_point=True, version="fake_act", calib_data=None,
                 split="train", text_column="text", duo_scaling=True, modules_to_not_convert=None,
                 export_compatible=False, quant_act=False, apply_clip=True, applyScale=True, samples=512,
                 n_parallel_calib_samples=None, max_chunk_memory=1024 * 1024 * 1024, quantUnet=True,
                 quantTextEncoder=False, quantVAE=False, quantTransformer=False, diffusion_model=True,
                 codeBookQuantInd=False, calibrate=False, pack_conv3x3=True, **unused_llm_flags) -> None:
        """Keyword names follow quantizer.py:36-81; `calibrate` switches the activation-aware search on for
        diffusion models (new); `pack_conv3x3` (version='gemm' only) stores the 3x3 / stride 1 resnet convolutions as
        packed int4 too (linear.QConv3x3, 4x smaller conv weights; the loop is ~3 % slower than with cuDNN on fp16
        fake-quant weights -- pass False to keep those).  LLM/VLM-only flags are accepted and ignored."""
        self.awq_model, self.model, self.tokenizer = awq_model, model, tokenizer
        self.w_bit, self.wv_bit, self.a_bit = w_bit, wv_bit, a_bit
        self.quantise_act = quantise_act
        self.weight_quant_conv_type, self.weight_quant_type = weight_quant_conv_type, weight_quant_type
        self.act_quant_conv_type, self.act_quant_conv_group_size = act_quant_conv_type, act_quant_conv_group_size
        self.group_size, self.zero_point, self.version = group_size, zero_point, version
        self.calib_data, self.split, self.text_column = calib_data, split, text_column
        self.duo_scaling = duo_scaling
        self.export_compatible, self.quant_act = export_compatible, quant_act
        self.apply_clip, self.applyScale = apply_clip, applyScale
        self.n_parallel_calib_samples, self.max_chunk_memory = n_parallel_calib_samples, max_chunk_memory
        self.modules_to_not_convert = modules_to_not_convert if modules_to_not_convert is not None else []
        self.samples = samples
        self.quantUnet, self.quantTextEncoder, self.quantVAE, self.quantTransformer = quantUnet, quantTextEncoder, quantVAE, quantTransformer
        self.codeBookQuantInd, self.diffusion_model = codeBookQuantInd, diffusion_model
        self.calibrate = calibrate
        self.pack_conv3x3 = pack_conv3x3
        self.search_log = []   # (block name, prev_op name, layer names, best ratio, best loss) per group
        if codeBookQuantInd:
            raise NotImplementedError("codebook quantisation is outside the quantized-linear hot path")
        if awq_model is not None:
            self.modules, self.module_kwargs, self.inps = self.init_quant()
        else:
            self.modules, self.module_kwargs, self.inps = {}, [], []

    MyTraversal = ModuleTraversal   # quantizer.py:142-159

    def _g(self, k):
        """group size used for a K-wide layer: the configured one, shrunk by 32 until it divides K exactly as
        the fake-quant path does (fake_quant.py:34-37; K = 320 -> 64).  The reference's AWQ branch would
        assert instead (quantizer.py:166); diffusion widths need the fallback."""
        from .fake_quant import _effective_group
        return _effective_group(k, self.group_size) if self.group_size > 0 else k

    # ------------------------------------------------------------------ quantizer.py:163-213
    def pseudo_quantize_tensor(self, w: torch.Tensor, bitWidth=4):
        """Per-group RTN; returns (w_fakequant, scales, zeros|None) exactly like quantizer.py:163-198."""
        org_w_shape = w.shape
        if self.group_size > 0:
            assert org_w_shape[-1] % self.group_size == 0
        else:
            assert w.dim() == 2
        dq, _, scales, zeros = ops.quant_group(w, self.group_size, bitWidth, zero_point=self.zero_point)
        return dq.reshape(org_w_shape), scales.view(org_w_shape[0], -1), (zeros.view(org_w_shape[0], -1) if zeros is not None else None)

    def pseudo_dequantize_tensor(self, w: nn.Linear, scales: torch.Tensor, zeros: Optional[torch.Tensor] = None):
        """quantizer.py:200-213 (plain tensor algebra on integer-valued weights)."""
        repeat_count = w.weight.data.shape[-1] // scales.shape[-1]
        scales = scales.repeat(1, repeat_count).reshape(w.weight.data.shape)
        if self.zero_point:
            zeros = zeros.repeat(1, repeat_count).reshape(w.weight.data.shape)
            return (w.weight.data - zeros) * scales
        return w.weight.data * scales

    # ------------------------------------------------------------------ init (quantizer.py:1049-1091, diffusion)
    def init_quant(self):
        modules = {"unet": [], "text_encoder": [], "vae": [], "transformer": []}
        if self.quantUnet:
            modules["unet"] = self.awq_model.get_model_layers_unet()
        if self.quantTextEncoder:
            modules["text_encoder"] = self.awq_model.get_model_layers_te()
        if self.quantVAE:
            modules["vae"] = self.awq_model.get_model_layers_vae()
        if self.quantTransformer:
            modules["transformer"] = self.awq_model.get_model_layers_transformers()
        return modules, [], []

    # ------------------------------------------------------------------ quantize (quantizer.py:386-425, diffusion)
    @torch.no_grad()
    def quantize(self, debugSavePath=None, debugPlot=False, shard=None):
        """Diffusion branch.  With `calibrate` the activation-aware search (quantizer.py:216-385 transplanted to
        UNet / MMDiT blocks) runs first: scales are folded (scale.py:37-84), clips applied (scale.py:25-34),
        and only then every Linear / Conv2d is swapped.  `shard` = (rank, world) restricts the search to this
        rank's blocks; dist.py gathers the results so every rank applies the identical list."""
        import time
        on_gpu = next(self.awq_model.denoiser().parameters()).is_cuda

        def now():
            if on_gpu:
                torch.cuda.synchronize()
            return time.perf_counter()
        if self.calibrate:
            results = self.search(shard=shard)
            t0 = now()
            self.apply_search_results(results)
            self.timings["apply_s"] = now() - t0
        t0 = now()
        for key in self.modules:
            for k, module_list in enumerate(self.modules[key]):
                root = self.awq_model.get_root(key, k)
                self.awq_model.set_quantized_components(f"{key}_{k + 1}" if (len(self.modules[key]) > 1 and k > 0) else key)
                for name, mod in module_list:
                    if next(mod.parameters(), None) is None:
                        continue
                    traversal = self.MyTraversal()
                    traversal.traverse(name, mod, root)
                    layers = [(parent, n, l) for parent, n, l in traversal.get_lin_conv()
                              if not any(key_ in n for key_ in self.modules_to_not_convert)]
                    if self.version in ("gemm", "w8a8"):
                        self._apply_quant_real(mod, layers, self.w_bit)
                    else:
                        self._apply_quant_fake_act(mod, layers, self.w_bit, debugStruct=None)
        if getattr(self, "timings", None) is not None:
            self.timings["swap_pack_s"] = now() - t0

    # ------------------------------------------------------------------ search over blocks (new orchestration)
    @torch.no_grad()
    def search(self, shard=None):
        """Returns {block_name: {"scales": [(prev_op_name, layer_names, scales)], "clip": [(layer_name, max_val)]}}
        for the blocks owned by this rank.  Blocks and their calibration inputs come from the adapter:
        `get_search_blocks()` -> {name: block}, `capture_block_inputs(names, shard, wanted_by)` -> {name: {linear_name: X}}
        (inputs of every Linear of the block from the FP passes over the calibration set, kept on the GPU).

        shard = (rank, world), SURVEY.md section 8(e):
          * the capture pass is data parallel over calibration batches (models.capture_block_inputs);
          * with >= 2 blocks per rank, blocks are assigned to ranks (dist.assign_blocks) and a rank runs both phases -- scale
            search over the full 20-point grid, then clip search on the scaled block -- for its blocks only;
          * with fewer blocks the RATIO GRID is split instead (dist.split_ratios: 3,3,3,3,2,2,2,2 at world 8): every rank
            evaluates its ratios for every group, ONE all_reduce(MIN) merges the [groups, 20] loss table, every rank takes
            the same first minimum (the reference's strict `<`, quantizer.py:739); the clip phase is then shared out block
            by block.
        Every loss is computed by one GPU from the same inputs in both modes, so the results do not depend on the world size."""
        import time
        from .dist import allreduce_min_losses, owners_of, split_ratios
        blocks = self.awq_model.get_search_blocks()
        names = list(blocks)
        rank, world = shard if shard is not None else (0, 1)
        ratios, wanted_by = None, None
        mine = clip_mine = names
        if world > 1:
            parts, ratio_split = owners_of(names, [self.awq_model.block_cost(blocks[n]) for n in names], world)
            if ratio_split:
                ratios = split_ratios(20, world)[rank]
                clip_mine = names[rank::world]
                wanted_by = {n: list(range(world)) for n in names}
            else:
                mine = clip_mine = parts[rank]
                wanted_by = {n: [r] for r, p in enumerate(parts) for n in p}
        tm = self.timings = {"capture_s": 0.0, "scale_search_s": 0.0, "clip_search_s": 0.0, "mode": "ratio-split" if ratios is not None else "blocks",
                             "blocks_searched": len(mine), "blocks_clipped": len(clip_mine)}
        on_gpu = next(self.awq_model.denoiser().parameters()).is_cuda

        def lap(key, t0):
            if on_gpu:
                torch.cuda.synchronize()
            tm[key] += time.perf_counter() - t0

        t0 = time.perf_counter()
        feats = self.awq_model.capture_block_inputs(mine, shard=shard, wanted_by=wanted_by)
        lap("capture_s", t0)
        tm.update(getattr(self.awq_model, "capture_timings", {}))
        # ---- phase A: scale search (loss tables), one collective when the ratio grid is split
        t0 = time.perf_counter()
        pending = {}
        for bname in mine:
            groups = self.awq_model.get_layers_for_scaling(blocks[bname], feats[bname]) if self.applyScale else []
            pending[bname] = [self._scale_losses(blocks[bname], ratios=ratios, **g) for g in groups]
        recs = [rec for bname in mine for rec in pending[bname]]
        if ratios is not None and recs:
            table = allreduce_min_losses(torch.stack([rec["losses"] for rec in recs]))
            for i, rec in enumerate(recs):
                rec["losses"] = table[i]
        scales = {bname: [self._finish_scale(rec) for rec in pending[bname]] for bname in mine}
        lap("scale_search_s", t0)
        # ---- phase B: clip search on the scaled block
        t0 = time.perf_counter()
        out = {}
        for bname in clip_mine:
            block, input_feat, scales_list = blocks[bname], feats[bname], [r[:3] for r in scales[bname]]
            res = {"scales": scales_list, "clip": []}
            if self.apply_clip:
                # the clip search runs on the SCALED weights and inputs (quantizer.py:312-336); do that on the block itself
                # and roll EVERY Linear back (apply_scale also folds into group members that are excluded from
                # quantisation by modules_to_not_convert), so that the gathered results are applied exactly once everywhere
                every = get_named_linears(block)
                backup = {n: l.weight.data.clone() for n, l in every.items()}
                prev_backup = self._snapshot_prev_ops(block, scales_list)
                apply_scale(block, scales_list, input_feat_dict=input_feat)
                named = exclude_layers_to_not_quantize(every, self.modules_to_not_convert)
                res["clip"] = self._search_best_clip(block, named, input_feat)
                for n, l in every.items():
                    l.weight.data = backup[n]
                self._restore_prev_ops(prev_backup)
            for prev, _, _, ratio, loss in scales[bname]:
                self.search_log.append((prev, ratio, loss))
            out[bname] = res
            del feats[bname]
        lap("clip_search_s", t0)
        return out

    def _snapshot_prev_ops(self, block, scales_list):
        from .scale import AdaLNShift, resolve_prev_op
        snap = []
        for prev_name, _, _ in scales_list:
            op = resolve_prev_op(block, prev_name)
            target = op.linear if isinstance(op, AdaLNShift) else op
            snap.append((target, {k: v.detach().clone() for k, v in target.state_dict().items()}))
        return snap

    @staticmethod
    def _restore_prev_ops(snap):
        for target, sd in snap:
            target.load_state_dict(sd)

    @torch.no_grad()
    def apply_search_results(self, results):
        blocks = self.awq_model.get_search_blocks()
        for bname, res in results.items():
            apply_scale(blocks[bname], res["scales"])
            apply_clip(blocks[bname], res["clip"])

    # ------------------------------------------------------------------ module swap (quantizer.py:491-577)
    def _apply_quant_fake_act(self, module, named_linears, bitWidth, debugStruct=None, debug=False):
        """quantizer.py:491-533 (diffusion branch): RTN fake-quant of every Linear / Conv2d, in place."""
        for parent, name, layer in named_linears:
            quantize_bmm_input = 'k_proj' in name or 'v_proj' in name or 'q_proj' in name
            if isinstance(layer, torch.nn.Linear):
                fake = WxAxLinear.from_float(layer, weight_quant=self.weight_quant_type, act_quant='per_token',
                                             quantize_output=quantize_bmm_input, n_bits_W=bitWidth, n_bits_A=self.a_bit,
                                             group_size_W=self.group_size, codeBookQuantInd=self.codeBookQuantInd)
            elif isinstance(layer, torch.nn.Conv2d):
                fake = WxAxConv2d.from_float(layer, weight_quant=self.weight_quant_conv_type, act_quant=self.act_quant_conv_type,
                                             quantize_output=self.quantise_act, act_group_size=self.act_quant_conv_group_size,
                                             n_bits_W=bitWidth, n_bits_A=self.a_bit, codeBookQuantInd=self.codeBookQuantInd)
            else:
                continue
            setattr(parent, name, fake)

    def _apply_quant_real(self, module, named_linears, bitWidth):
        """version == 'gemm': the reference's `_apply_quant` (quantizer.py:535-577) with the real packed module --
        pseudo_quantize_tensor + transpose + WQLinear_GEMM.from_linear fused into one RTN+pack kernel.
        version == 'w8a8': int8 per-channel weights + per-token activations (fake_quant.py:86-93,109-118).
        1x1 Conv2d layers become QConv1x1 (the same kernels on the token view), 3x3 / stride 1 / pad 1 layers QConv3x3
        (kernel c as an implicit GEMM); other convolutions keep the fake-quant path."""
        from .fake_quant import _effective_group
        from .linear import (QConv1x1, QConv3x3, W8A8Linear, conv_group, is_conv3x3_gemm, is_pointwise_conv, w4a16_kernel_ok,
                             w8a8_kernel_ok)
        for parent, name, layer in named_linears:
            if isinstance(layer, torch.nn.Linear):
                if self.version == "w8a8":
                    if not w8a8_kernel_ok(layer.in_features, layer.out_features):   # shapes kernel (d) does not tile stay fake-quant
                        self._apply_quant_fake_act(module, [(parent, name, layer)], 8)
                        continue
                    new = W8A8Linear.from_float(layer)
                else:
                    g = _effective_group(layer.in_features, self.group_size) if self.group_size > 0 else layer.in_features
                    if not w4a16_kernel_ok(layer.in_features, layer.out_features, g):   # e.g. group_size = -1 on K = 320
                        self._apply_quant_fake_act(module, [(parent, name, layer)], bitWidth)
                        continue
                    new = WQLinear_GEMM.from_linear(layer, bitWidth, g)
                setattr(parent, name, new)
            elif isinstance(layer, torch.nn.Conv2d):
                # pointwise convolutions are GEMMs over channels: real packed modules on the token view
                # (SURVEY.md section 8(f) row 3); everything else keeps the fake-quant weights + cuDNN.
                g = _effective_group(layer.in_channels, self.group_size) if self.group_size > 0 else layer.in_channels
                if is_pointwise_conv(layer) and layer.out_channels % 8 == 0 and layer.weight.dtype != torch.float32:
                    if self.version == "w8a8" and w8a8_kernel_ok(layer.in_channels, layer.out_channels):
                        setattr(parent, name, QConv1x1.from_conv_w8a8(layer))
                        continue
                    if self.version != "w8a8" and w4a16_kernel_ok(layer.in_channels, layer.out_channels, g):
                        setattr(parent, name, QConv1x1.from_conv_w4a16(layer, bitWidth, g))
                        continue
                # 3x3 / stride 1 / pad 1: packed int4 weights on the implicit-GEMM form of kernel (c)
                g3 = conv_group(9 * layer.in_channels, self.group_size)
                if (self.version == "gemm" and self.pack_conv3x3 and is_conv3x3_gemm(layer) and g3
                        and layer.weight.dtype != torch.float32):
                    setattr(parent, name, QConv3x3.from_conv(layer, bitWidth, g3))
                    continue
                self._apply_quant_fake_act(module, [(parent, name, layer)], 8 if self.version == "w8a8" else bitWidth)

    def _apply_quant(self, module, named_linears: Dict[str, nn.Linear], bitWidth):
        """quantizer.py:535-577 for a dict of named Linears (LLM-style call shape)."""
        for name, linear_layer in named_linears.items():
            if self.version != "gemm":
                raise ValueError(f"Unknown version {self.version}")
            q_linear = WQLinear_GEMM.from_linear(linear_layer, bitWidth, self.group_size)
            set_op_by_name(module, name, q_linear)

    # ------------------------------------------------------------------ quantizer.py:580-603
    @torch.no_grad()
    def _module_forward(self, x: torch.Tensor, module: torch.nn.Module, module_kwargs: Dict) -> torch.Tensor:
        if self.n_parallel_calib_samples is None:
            out = module(x, **module_kwargs)
            return out[0] if isinstance(out, tuple) else out
        outs = []
        for xp in torch.split(x, self.n_parallel_calib_samples):
            o = module(xp, **module_kwargs)
            outs.append(o[0] if isinstance(o, tuple) else o)
        return torch.cat(outs, dim=0)

    # ------------------------------------------------------------------ quantizer.py:606-676
    @torch.no_grad()
    def _scale_losses(self, module, prev_op, layers: List[nn.Linear], inp: torch.Tensor, module2inspect=None, kwargs={}, ratios=None):
        """Steps 1-4 of `_search_best_scale` up to the loss table: returns the group's record {prev, names, x_mean, w_mean,
        losses [20] float64 on the device (inf where `ratios` left a grid point out)}."""
        if module2inspect is None:
            assert len(layers) == 1
            module2inspect = layers[0]
        kwargs = {k: v for k, v in kwargs.items() if k != "use_cache"}
        inp = inp.to(next(module2inspect.parameters()).device)
        # [STEP 1] per-channel mean of the group-normalised weights (quantizer.py:627-637): fused kernel, fp32 sums
        weight = torch.cat([m.weight for m in layers], dim=0)
        w_mean = (ops.awq_wsum(weight, self._g(weight.shape[1])) / weight.shape[0]).to(weight.dtype)
        # [STEP 2] per-channel mean |x| in fp32 (quantizer.py:642-659): one reduction kernel, no CPU round trip
        n_tok = inp.numel() // inp.shape[-1]
        x_mean = (ops.colabssum(inp) / n_tok).to(inp.dtype)
        # [STEP 3] reference output
        fp16_output = self._module_forward(inp, module2inspect, kwargs)
        # [STEP 4] grid search
        losses = self._grid_losses(inp, w_mean, x_mean, module2inspect, layers, fp16_output, kwargs, ratios)
        from .scale import describe_prev_op
        return {"prev": describe_prev_op(module, prev_op), "names": tuple(get_op_name(module, m) for m in layers),
                "x_mean": x_mean, "w_mean": w_mean, "losses": losses}

    def _finish_scale(self, rec, n_grid=20):
        """argmin of a (complete) loss table -> (prev_op_name, layer_names, best_scales, best_ratio, best_loss)."""
        losses = rec["losses"]
        best = int(torch.argmin(losses).item())   # first minimum == the reference's strict `<` (quantizer.py:739)
        if not torch.isfinite(losses[best]):
            logging.debug(losses.tolist())
            raise Exception
        best_scales = self._ratio_scales(rec["x_mean"].view(-1), rec["w_mean"].view(-1), best / n_grid)
        assert torch.isnan(best_scales).sum() == 0, best_scales
        self.last_losses, self.last_best_ratio = losses, best / n_grid
        return (rec["prev"], rec["names"], best_scales.detach(), best / n_grid, float(losses[best].item()))

    @torch.no_grad()
    def _search_best_scale(self, module, prev_op, layers: List[nn.Linear], inp: torch.Tensor, module2inspect=None, kwargs={}):
        prev, names, best_scales, ratio, loss = self._finish_scale(self._scale_losses(module, prev_op, layers, inp, module2inspect, kwargs))
        self.search_log.append((prev, ratio, loss))
        return (prev, names, best_scales)

    # ------------------------------------------------------------------ quantizer.py:678-751
    def _ratio_scales(self, x_mean, w_mean, ratio):
        """quantizer.py:717-725: K-length vector algebra, left in torch on the device."""
        if self.duo_scaling:
            scales = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
        else:
            scales = x_mean.pow(ratio).clamp(min=1e-4).view(-1)
        scales = scales / (scales.max() * scales.min()).sqrt()
        scales[torch.isinf(scales)] = 1
        scales[torch.isnan(scales)] = 1
        return scales

    def _ratio_scales_all(self, x_mean, w_mean, ratios, n_grid=20):
        """The scale vectors of several grid points at once, [len(ratios), K]: row j is bit-identical to
        `_ratio_scales(x_mean, w_mean, ratios[j] / n_grid)` (the two `pow` keep their Python-scalar exponents -- torch
        special-cases 0, 0.5, 1 -- everything after them is the same elementwise / row-wise-exact op applied to all rows),
        with ~12 launches for the whole grid instead of ~13 per grid point: the search loop is host-bound."""
        if self.duo_scaling:
            num = torch.stack([x_mean.pow(i / n_grid) for i in ratios])
            den = torch.stack([w_mean.pow(1 - i / n_grid) for i in ratios])
            scales = (num / (den + 1e-4)).clamp(min=1e-4)
        else:
            scales = torch.stack([x_mean.pow(i / n_grid) for i in ratios]).clamp(min=1e-4)
        scales = scales / (scales.amax(dim=1, keepdim=True) * scales.amin(dim=1, keepdim=True)).sqrt()
        scales.masked_fill_(torch.isinf(scales), 1)
        scales.masked_fill_(torch.isnan(scales), 1)
        return scales

    @torch.no_grad()
    def _grid_losses(self, x, w_mean, x_mean, module2inspect, linears2scale: List[nn.Linear], fp16_output, kwargs: Dict = {},
                     ratios=None, n_grid=20):
        """L(s) = || Q(W * s) (s^-1 * X) - W * X || for the grid points in `ratios` (default: all 20) -> float64 [20]
        on the device, inf elsewhere and where the loss is NaN."""
        ratios = list(range(n_grid)) if ratios is None else list(ratios)
        x_mean, w_mean = x_mean.view(-1), w_mean.view(-1)
        org = [fc.weight.data for fc in linears2scale]
        scratch = [torch.empty_like(w) for w in org]
        losses = torch.full((n_grid,), float("inf"), dtype=torch.float64, device=x.device)
        g = self._g(org[0].shape[1])
        s_all = self._ratio_scales_all(x_mean, w_mean, ratios, n_grid).to(org[0].dtype)
        try:
            for j, i in enumerate(ratios):
                s_w = s_all[j]
                for fc, w, buf in zip(linears2scale, org, scratch):
                    # Q(W * s) / s in one kernel (quantizer.py:727-730)
                    ops.quant_group(w, g, 4, zero_point=self.zero_point, pre_mul=s_w, post_div=s_w, want_scales=False, out=buf)
                    fc.weight.data = buf
                int_w_output = self._module_forward(x, module2inspect, kwargs)
                ops.sqdiff_sum(fp16_output, int_w_output, out=losses[i])          # quantizer.py:754-783, the sum ...
        finally:
            for fc, w in zip(linears2scale, org):
                fc.weight.data = w
        losses = losses / fp16_output.numel()                                      # ... and the mean, once for the table
        return torch.where(torch.isnan(losses), torch.full_like(losses, float("inf")), losses)

    @torch.no_grad()
    def _compute_best_scale(self, x, w_mean, x_mean, module2inspect, linears2scale: List[nn.Linear], fp16_output, kwargs: Dict = {},
                            ratios=None):
        """quantizer.py:678-751: returns best_scales [K] on the device; the loss vector is kept in `self.last_losses`."""
        losses = self._grid_losses(x, w_mean, x_mean, module2inspect, linears2scale, fp16_output, kwargs, ratios)
        return self._finish_scale({"prev": None, "names": (), "x_mean": x_mean, "w_mean": w_mean, "losses": losses})[2]

    @torch.no_grad()
    def _compute_loss(self, fp16_output, int_w_output, device=None):
        """quantizer.py:754-783 -> python float."""
        return (ops.sqdiff_sum(fp16_output, int_w_output) / fp16_output.numel()).item()

    # ------------------------------------------------------------------ quantizer.py:786-863
    @torch.no_grad()
    def _search_best_clip(self, layer, named_linears, input_feat):
        clip_list = []
        avoid_clipping = ["q_", "k_", "query", "key", "Wqkv", "to_q", "to_k", "add_q_proj", "add_k_proj"]
        for name in named_linears:
            if any(a in name for a in avoid_clipping):   # inputs of the QK^T bmm are hard to clip precisely
                continue
            if name not in input_feat:
                continue
            clip_list.append((name, self._compute_best_clip(named_linears[name].weight, input_feat[name])))
        return clip_list

    @torch.no_grad()
    def _compute_best_clip(self, w: torch.Tensor, input_feat: torch.Tensor, n_grid=20, max_shrink=0.5, n_sample_token=512):
        """quantizer.py:805-863.  Groups of 64 / 128 in fp16 / bf16 (every diffusion layer) run the clip-search kernel.
        Other group sizes keep the batched-GEMM form below: the per-group dot products sum_g x*w as one batched GEMM per
        out-row batch ([G] x [co_b, g] x [g, n_tok]) instead of a co_b x n_tok x K broadcast product, Q(clamp(w)) by the
        fused quantise kernel with `clip_max`.  The reference rounds every x*w product to fp16 before summing; both forms
        keep fp32 products, so err values agree to ~1e-3 relative and best_max can differ on near-ties."""
        assert w.dim() == 2
        co, ci = w.shape
        gs = self._g(ci)
        G = ci // gs
        x = input_feat.view(-1, input_feat.shape[-1])
        x = x[:: max(1, x.shape[0] // n_sample_token)]
        if gs in (64, 128) and w.dtype in (torch.float16, torch.bfloat16) and x.dtype == w.dtype:
            # the search kernel: err = d^T C d with the group's Gram matrix C, shrink levels
            # quantised on the fly, no [G, co, n_tok] temporaries (include/qdm.h: qdm_awq_clip_search)
            return ops.awq_clip_search(w, x, gs, 4, self.zero_point, n_grid, max_shrink)
        xg = x.reshape(-1, G, gs).permute(1, 2, 0).contiguous().float()      # [G, g, n_tok]
        # The reference walks out-rows in batches of 256 / 64 because its broadcast product needs
        # co_b x n_tok x K temporaries (quantizer.py:827); rows are independent, and the batched-GEMM form only
        # needs G x co_b x n_tok fp32, so the batch is as large as ~1 GB of temporaries allows.
        assert co % (256 if co % 256 == 0 else 64) == 0
        n_tok = xg.shape[2]
        oc_batch = co
        while oc_batch * G * n_tok * 4 > (1 << 30) and oc_batch % 2 == 0 and oc_batch > 64:
            oc_batch //= 2
        best_all = []
        for b in range(co // oc_batch):
            wb = w[b * oc_batch:(b + 1) * oc_batch].contiguous()             # [co_b, ci]
            org_max = ops.rowabsmax(wb.reshape(-1, gs)).reshape(oc_batch, G)  # exact
            best_max = org_max.clone()
            min_errs = torch.full((oc_batch, G), 1e9, dtype=torch.float32, device=w.device)
            org_out = torch.bmm(wb.reshape(oc_batch, G, gs).permute(1, 0, 2).float(), xg)   # [G, co_b, n_tok]
            for i_s in range(int(max_shrink * n_grid)):
                max_val = org_max * (1 - i_s / n_grid)
                q_w = ops.quant_group(wb, gs, 4, zero_point=self.zero_point, clip_max=max_val.reshape(-1), want_scales=False)[0]
                cur_out = torch.bmm(q_w.reshape(oc_batch, G, gs).permute(1, 0, 2).float(), xg)
                err = (cur_out - org_out).pow(2).mean(dim=2).t()             # [co_b, G]
                better = err < min_errs
                min_errs[better] = err[better]
                best_max[better] = max_val[better]
            best_all.append(best_max)
        return torch.cat(best_all, dim=0).unsqueeze(-1)                      # [co, G, 1]
