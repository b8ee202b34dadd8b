"""Host-side mirror of the reference's quantize/quantizer_SQ.py (class SqQuantizer), diffusion branch:
absmax hooks -> smoothing scale act^alpha / w^(1-alpha) folded into (LayerNorm, following Linears) -> RTN swap.

Deviations: `alpha` is a parameter (the reference hard-codes 0.80 for diffusion at
quantizer_SQ.py:349 while BASELINE config 3 asks for 0.5); the NameError at :386 is not reproduced; adapters
other than SD1.x provide smoothing blocks too; `version='w8a8'` swaps in real int8 modules (kernel d).
"""
import torch

from . import ops
from .calib_data import Fused_Mean_Max_Activation_Hook, Mean_Max_Activation_Hook, apply_hook, get_calib_dataset_dm, run_calibration
from .quantizer import AwqQuantizer


class SqQuantizer(AwqQuantizer):
    def __init__(self, awq_model, *args, alpha=0.5, calib_n_samples=96, calib_batch_size=8, calib_num_infer_steps=50,
                 calib_prompts="clip-benchmark/wds_mscoco_captions2017", fused_stats=False, **kwargs):
        """alpha / calibration-set knobs are explicit; everything else follows quantizer_SQ.py:35-140."""
        self.alpha = alpha
        self.calib_n_samples, self.calib_batch_size = calib_n_samples, calib_batch_size
        self.calib_num_infer_steps, self.calib_prompts = calib_num_infer_steps, calib_prompts
        self.smooth_log = {}
        self.fused_stats = fused_stats   # one-pass in-place hook statistic instead of per-call tensors
        super().__init__(awq_model, *args, **kwargs)

    # ------------------------------------------------------------------ quantizer_SQ.py:396-431
    @torch.no_grad()
    def smooth_ln_fcs(self, ln, fcs, act_scales, model_type="transformers", alpha=0.5):
        if not isinstance(fcs, list):
            fcs = [fcs]
        for fc in fcs:
            assert ln.weight.numel() == fc.in_features == act_scales.numel()
        device, dtype = fcs[0].weight.device, fcs[0].weight.dtype
        act_scales = act_scales.to(device=device, dtype=dtype)
        # max over the fcs of the per-input-channel |W| max: column reduction kernel in running-max mode
        weight_scales = ops.colabsmax(fcs[0].weight.data)
        for fc in fcs[1:]:
            ops.colabsmax(fc.weight.data, out=weight_scales, running=True)
        weight_scales = weight_scales.clamp(min=1e-5)
        scales = (act_scales.pow(alpha) / weight_scales.pow(1 - alpha)).clamp(min=1e-5).to(device).to(dtype)
        ln.weight.div_(scales)
        if hasattr(ln, "bias") and ln.bias is not None:
            ln.bias.div_(scales)
        for fc in fcs:
            fc.weight.mul_(scales.view(1, -1))
        return scales

    def apply_hooks_to_smoothing_blocks(self, blocks):
        """quantizer_SQ.py:1064-1070."""
        hook_cls = Fused_Mean_Max_Activation_Hook if self.fused_stats else Mean_Max_Activation_Hook
        return {name: apply_hook(block, hook_cls) for name, block in blocks.items()}

    # ------------------------------------------------------------------ quantizer_SQ.py:323-391
    @torch.no_grad()
    def quantize(self, debugSavePath=None, debugPlot=False, samples=None, shard=None):
        """shard = (rank, world): the calibration batches are run data parallel (batch i on rank i % world) with the fused
        one-pass hooks, and ONE exact fp64 all_reduce of their accumulators (dist.allreduce_hook_stats) gives every rank
        the statistic of the whole calibration set: smoothing scales and codes are identical for every world size."""
        rank, world = shard if shard is not None else (0, 1)
        if world > 1:
            self.fused_stats = True   # per-call tensors cannot be merged across ranks; the fp64 accumulators can, exactly
        smoothing_blocks = self.awq_model.get_smoothing_blocks()
        hook_d = self.apply_hooks_to_smoothing_blocks(smoothing_blocks)
        pipe = self.awq_model.get_pipeline()
        if samples is None:
            samples = get_calib_dataset_dm(model_pipeline=pipe, text_dataset=self.calib_prompts,
                                           batch_size=self.calib_batch_size, n_samples=self.calib_n_samples, seed=42,
                                           device=pipe.device, split="test", text_column="txt")
        run_calibration(pipe, samples[rank::world] if world > 1 else samples, None, self.calib_num_infer_steps)
        if world > 1:
            from .dist import allreduce_hook_stats
            allreduce_hook_stats(hook_d)
        for block_name, block_module in smoothing_blocks.items():
            for group in self.awq_model.get_layers_for_scaling_unet(block_module, hook_d[block_name]):
                s = self.smooth_ln_fcs(group['prev_op'], group['layers'], group['activations_max'][0], alpha=self.alpha)
                self.smooth_log.setdefault(block_name, []).append(s)
            for _, hook in hook_d[block_name].items():
                hook.clear()
                hook.hook_handle.remove()
        hook_d.clear()
        # RTN swap of every Linear / Conv (quantizer_SQ.py:358-391)
        self.calibrate = False
        super().quantize(debugSavePath, debugPlot)
