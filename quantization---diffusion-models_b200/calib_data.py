"""Calibration plumbing with the semantics of the reference's utils/calib_data.py (diffusion part):
per-Linear activation hooks, deterministic calibration latents (seed 42) and the pipeline run loop.
The per-call statistic is computed by the col-absmax reduction kernel (kernel a)."""
import torch

from . import ops


class Mean_Max_Activation_Hook:
    """utils/calib_data.py:105-124: per forward call, max over tokens of |x| per input channel.
    `max_scales[step]` keeps one [C] vector per call exactly like the reference, so that
    `mean_of_dict` (models/StableDiffusion1_x.py:104-112) reduces the same values in the same way."""

    def __init__(self):
        self.hook_handle = None
        self.max_scales = {}
        self.step = 0

    def __call__(self, module, module_in, module_out):
        self.max_scales[self.step] = ops.colabsmax(module_in[0])
        self.step += 1

    def clear(self):
        self.max_scales = []


class AccumulatedMax:
    """What `Fused_Mean_Max_Activation_Hook.max_scales` hands to `mean_of_dict`: the fp64 sum over calls of the
    per-call column maxima plus the call count (instead of one retained [C] tensor per call)."""

    def __init__(self, acc_maxsum, n_calls, dtype):
        self.acc_maxsum, self.n_calls, self.dtype = acc_maxsum, n_calls, dtype

    def __len__(self):
        return self.n_calls

    def mean_over_calls(self):
        return (self.acc_maxsum / self.n_calls).to(self.dtype)


class Fused_Mean_Max_Activation_Hook:
    """the statistic of Mean_Max_Activation_Hook (utils/calib_data.py:105-124, mean over
    calls of the per-call column |x| max) collected by ONE pass of the one-pass hook kernel (`qdm_colstats`) per
    call, folded in place into fp64 accumulators on the GPU: no `max_scales[step]` tensors are retained, and the
    same pass also maintains the running max (quantizer_SQ.py:1077-1084) and -- with `want_abssum` -- the
    numerator of AWQ's x_mean (quantizer.py:642-659).  The fp64 sum of fp16 maxima is exact, so the mean is the
    correctly rounded one: identical to the reference's `torch.mean(torch.stack(...))` except where the
    reference's own fp32 summation order lands across a rounding boundary (<= 1 ulp; tests bound it)."""

    def __init__(self, want_abssum=False):
        self.hook_handle = None
        self.want_abssum = want_abssum
        self.acc_maxsum = None
        self.acc_abssum = None
        self.running_max = None
        self.dtype = None
        self.step = 0
        self.rows = 0

    def __call__(self, module, module_in, module_out):
        x = module_in[0]
        c = x.shape[-1]
        first = self.acc_maxsum is None
        if first:
            self.dtype = x.dtype
            self.acc_maxsum = torch.zeros(c, dtype=torch.float64, device=x.device)
            self.running_max = torch.empty(c, dtype=x.dtype, device=x.device)
            if self.want_abssum:
                self.acc_abssum = torch.zeros(c, dtype=torch.float64, device=x.device)
        ops.colstats(x, out_max=self.running_max, running=not first, acc_maxsum=self.acc_maxsum,
                     acc_abssum=self.acc_abssum)
        self.step += 1
        self.rows += x.numel() // c

    @property
    def max_scales(self):
        return AccumulatedMax(self.acc_maxsum, self.step, self.dtype)

    def x_mean(self):
        """mean over all tokens seen of |x| per channel, in the activation dtype (quantizer.py:652-659)."""
        if self.acc_abssum is None:
            raise RuntimeError("the hook was created without want_abssum=True")
        return (self.acc_abssum / self.rows).to(self.dtype)

    def clear(self):
        self.acc_maxsum = self.acc_abssum = self.running_max = None
        self.step = self.rows = 0


class Running_Max_Activation_Hook:
    """LLM-style statistic of quantize/quantizer_SQ.py:1077-1084: running max over calls, folded in place by
    the kernel's running mode (no per-call tensors retained)."""

    def __init__(self):
        self.hook_handle = None
        self.running = None
        self.step = 0

    def __call__(self, module, module_in, module_out):
        x = module_in[0]
        if self.running is None:
            self.running = ops.colabsmax(x)
        else:
            ops.colabsmax(x, out=self.running, running=True)
        self.step += 1


class Input_Capture_Hook:
    """Keeps the inputs of a Linear on the GPU (the reference caches them on the CPU, quantizer.py:1097-1100).
    `per_call` bounds the rows kept per forward call (evenly strided subsample, copied, so the activation itself can be
    freed): with per_call = max_tokens / total calls the memory of a capture pass is bounded by max_tokens rows per
    Linear however long the calibration runs.  Chunks carry a global call id (`next_id`, set by the capture loop per
    calibration batch) so that captures made data parallel can be merged in single-process order (dist.exchange_captures)."""

    def __init__(self, max_tokens=None, per_call=None, arena=None):
        self.hook_handle = None
        self.chunks = []          # [(call_id, X[rows, K])]
        self.max_tokens = max_tokens
        self.per_call = per_call
        self.next_id = 0
        # optional pre-allocated [rows, K] buffer (a slice of ONE arena per capture pass, models.capture_block_inputs): the
        # kept rows are copied into consecutive slices of it instead of a fresh allocation per call -- thousands of small
        # allocations interleaved with the forward's temporaries cost ~2 s of cudaMalloc on the 38-block SD3.5-L pass
        self.arena = arena
        self.used = 0

    def __call__(self, module, module_in, module_out):
        x = module_in[0].detach()
        x = x.reshape(-1, x.shape[-1])
        if self.per_call is not None and x.shape[0] > self.per_call:
            x = x[:: x.shape[0] // self.per_call][: self.per_call]
        if self.arena is not None and self.used + x.shape[0] <= self.arena.shape[0] and x.dtype == self.arena.dtype:
            dst = self.arena[self.used:self.used + x.shape[0]]
            dst.copy_(x)
            self.used += x.shape[0]
            self.chunks.append((self.next_id, dst))
        else:
            self.chunks.append((self.next_id, x.clone()))
        self.next_id += 1

    @staticmethod
    def merge(chunks, max_tokens=None):
        cs = [c for _, c in sorted(chunks, key=lambda t: t[0])]
        adjacent = all(a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype and a.shape[1] == b.shape[1] and
                       a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() and
                       a.storage_offset() + a.numel() == b.storage_offset() for a, b in zip(cs, cs[1:]))
        if len(cs) > 1 and adjacent:      # consecutive slices of one arena: the concatenation already exists
            x = cs[0].new_empty(0).set_(cs[0].untyped_storage(), cs[0].storage_offset(), (sum(c.shape[0] for c in cs), cs[0].shape[1]))
        else:
            x = torch.cat(cs, dim=0) if len(cs) > 1 else cs[0]
        if max_tokens is not None and x.shape[0] > max_tokens:
            x = x[:: x.shape[0] // max_tokens][:max_tokens].contiguous()
        return x

    def cat(self):
        return self.merge(self.chunks, self.max_tokens)


def apply_hook(module: torch.nn.Module, hook_cls=Mean_Max_Activation_Hook):
    """utils/calib_data.py:216-224: one hook per nn.Linear, keyed by the name relative to `module`."""
    hook_dictionary = {}
    for name, submod in module.named_modules():
        if isinstance(submod, torch.nn.Linear):
            hook = hook_cls()
            hook.hook_handle = submod.register_forward_hook(hook)
            hook_dictionary[name] = hook
    return hook_dictionary


def remove_hooks(hook_d):
    """utils/calib_data.py:247-249 (the reference iterates the dict itself and fails; `.items()` is meant)."""
    for _, v in hook_d.items():
        v.hook_handle.remove()


def generate_latents(pipe, batch_size, device, generator):
    """utils/calib_data.py:139-171: N(0,1) latents of the denoiser's input shape in the pipeline dtype."""
    shape = (batch_size, pipe.latent_channels, pipe.latent_size, pipe.latent_size)
    return torch.randn(shape, generator=generator, dtype=torch.float32).to(device=device, dtype=pipe.dtype)


def get_calib_dataset_dm(model_pipeline, text_dataset, batch_size=4, n_samples=100, seed=42, device="cuda",
                         split=None, text_column=None, cut_off_captions=200):
    """utils/calib_data.py:174-213.  There is no network here, so `text_dataset` must be a list of prompts
    (any hashable) -- the skeleton pipelines turn a prompt into a deterministic synthetic embedding."""
    assert n_samples % batch_size == 0, "The batch_size, doesnt divide the dataset, choose an appropriate batch_size"
    if isinstance(text_dataset, str):
        prompts = [f"{text_dataset}#{i}" for i in range(n_samples)]   # stand-in captions (offline)
    else:
        prompts = list(text_dataset)
    generator = torch.Generator().manual_seed(seed)
    calib_data = []
    for i in range(n_samples // batch_size):
        prompt_batch = prompts[i * batch_size:(i + 1) * batch_size]
        calib_data.append((prompt_batch, generate_latents(model_pipeline, batch_size, device, generator)))
    return calib_data


def run_calibration(pipeline, samples, callback=None, n_inference_steps=50, cfg=7.5):
    """utils/calib_data.py:227-245."""
    for i, sample in enumerate(samples):
        if callback is not None:
            callback.set_batch_num(i)
        pipeline(prompt=sample[0], latents=sample[1], callback_on_step_end=callback,
                 num_inference_steps=n_inference_steps, guidance_scale=cfg, num_images_per_prompt=1)
