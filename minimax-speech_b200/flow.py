"""Drop-in replacements for the reference's flow-matching decoder classes.

Same constructor / ``forward`` signatures, same ``state_dict`` key schema (reference checkpoints load
with ``load_state_dict`` unchanged) as

  * ``cosyvoice.flow.decoder.CausalConditionalDecoder``      speech/cosyvoice/flow/decoder.py:294-496
  * ``cosyvoice.flow.flow_matching.ConditionalCFM``           speech/cosyvoice/flow/flow_matching.py:21-155
  * ``cosyvoice.flow.flow_matching.CausalConditionalCFM``     speech/cosyvoice/flow/flow_matching.py:317-348

so ``speech/config.yaml:89,105`` can name these classes instead (``!new:minimax_speech_b200.flow....``).
All arithmetic runs in the CUDA library behind include/ls_b200.h; these modules only hold parameters,
validate arguments and marshal pointers.  Inference only: ``compute_loss*`` (training) is out of scope.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import native, ops, synth


class _Holder(nn.Module):
    """Parameter container addressed by the reference's dotted state_dict names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder")


def _register_tree(root, state_dict):
    for name, value in state_dict.items():
        mod = root
        *path, leaf = name.split(".")
        for part in path:
            if not hasattr(mod, part):
                mod.add_module(part, _Holder())
            mod = getattr(mod, part)
        mod.register_parameter(leaf, nn.Parameter(value.clone(), requires_grad=False))


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _as_f32(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


def _check_prefix_mask(mask, sync_ok=False):
    """The kernels take per-utterance lengths; the reference always builds prefix masks
    (``~make_pad_mask(len)``, flow.py:478,493).  A CPU mask is checked here.  A CUDA mask is checked on the device by the
    tensor-core engine itself (mask_to_lengths_kernel raises a sticky flag that the next call on the handle reports):
    reading the verdict back here would synchronise the host with the stream at the top of every call.  ``sync_ok``
    (the fp32 validation mode) checks CUDA masks here anyway."""
    if mask.is_cuda and not sync_ok:
        return
    m = mask != 0
    if bool((m[..., 1:] & ~m[..., :-1]).any()):
        raise ValueError("mask must be a prefix (right-padding) mask")


@torch.compiler.assume_constant_result
def _t_span_values(n_timesteps, scheduler):
    """The Euler time grid as a tuple of Python floats holding the reference's fp32 values (flow_matching.py:56-58 /
    340-342: ``linspace`` then ``1 - cos(t * 0.5 * pi)``, all in fp32).  Constant for given arguments, so
    torch.compile folds it."""
    t_span = torch.linspace(0, 1, n_timesteps + 1, dtype=torch.float32)
    if scheduler == "cosine":
        t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
    return tuple(float(v) for v in t_span)


class CausalConditionalDecoder(nn.Module):
    """Estimator.  ``forward(x, mask, mu, t, spks, cond, streaming)`` -> ``[rows, out_channels, T]``."""

    def __init__(self, in_channels=320, out_channels=80, channels=(256,), dropout=0.0, attention_head_dim=64,
                 n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn="gelu", static_chunk_size=50,
                 num_decoding_left_chunks=-1, weight_seed=1986, precision="bf16", _causal=True):
        super().__init__()
        self.precision = native.check_precision(precision, fp16_ok=True)
        channels = tuple(channels)
        if channels != (256,) or attention_head_dim != 64 or act_fn != "gelu" or in_channels != 4 * out_channels:
            raise NotImplementedError("B200 estimator covers config.yaml's CausalConditionalDecoder: "
                                      "channels=[256], head_dim 64, act_fn gelu, in_channels = 4*out_channels")
        if static_chunk_size != 50 or num_decoding_left_chunks != -1:
            raise NotImplementedError("streaming mask: static_chunk_size=50 with all left chunks")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.static_chunk_size, self.num_decoding_left_chunks = static_chunk_size, num_decoding_left_chunks
        _register_tree(self, synth.estimator_state_dict(
            weight_seed, "reference", in_channels=in_channels, out_channels=out_channels, channels=256,
            n_blocks=n_blocks, num_mid_blocks=num_mid_blocks, num_heads=num_heads, head_dim=attention_head_dim,
            causal=_causal))
        self._handle = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    def invalidate(self):
        self._handle = None

    def handle(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the B200 hot path runs on CUDA tensors only (no CPU fallback)")
        if self._handle is None or self._handle.device != device or self._handle.precision != self.precision:
            self._handle = native.FlowHandle(self.state_dict(), device, self.precision)
        return self._handle

    @torch.inference_mode()
    def forward(self, x, mask, mu, t, spks=None, cond=None, streaming=False):
        return self.run(x, mask, mu, t, spks, cond, streaming)

    def run(self, x, mask, mu, t, spks=None, cond=None, streaming=False):
        """``forward`` without the inference_mode decorator (the form torch.compile traces)."""
        dev = x.device
        if x.dim() != 3 or x.shape[1] != self.out_channels or x.shape[2] < 1:
            raise ValueError(f"x must be [rows, {self.out_channels}, T >= 1], got {tuple(x.shape)}")
        rows, F, T = x.shape
        spks = torch.zeros(rows, self.out_channels, device=dev) if spks is None else spks
        cond = torch.zeros_like(x) if cond is None else cond
        if tuple(mu.shape) != (rows, F, T) or tuple(cond.shape) != (rows, F, T) or tuple(mask.shape) != (rows, 1, T) \
                or tuple(spks.shape) != (rows, F):
            raise ValueError("estimator inputs: x, mu, cond [rows,80,T], mask [rows,1,T], spks [rows,80]")
        _check_prefix_mask(mask, self.precision == "fp32")
        t = t.reshape(-1).expand(rows) if t.numel() == 1 else t
        if t.numel() != rows:
            raise ValueError(f"t must have {rows} entries")
        out = torch.ops.ls_b200.estimator_forward(self.handle(dev).key, _as_f32(x, dev), _as_f32(mask, dev),
                                                  _as_f32(mu, dev), _as_f32(t, dev), _as_f32(spks, dev),
                                                  _as_f32(cond, dev), bool(streaming))
        return out.to(x.dtype)


class ConditionalDecoder(CausalConditionalDecoder):
    """The non-causal estimator, ``cosyvoice.flow.decoder.ConditionalDecoder`` (speech/cosyvoice/flow/decoder.py:88-291): the
    same U-Net with matcha's ``Block1D`` blocks -- Conv1d(k 3, padding 1) -> GroupNorm(8) -> Mish -- in place of the causal
    conv + LayerNorm ones, and full (never block-causal) attention.  Same ``forward(x, mask, mu, t, spks, cond)``, same
    state_dict key schema as the reference class (GroupNorm parameters at ``block.1``).  GroupNorm statistics are taken over
    each utterance's own frames (one reference call per utterance).  Covered geometry: ``channels=[256]`` (no down/up-sampling
    level), 8 heads x 64, ``act_fn='gelu'`` -- config.yaml's estimator with the non-causal class swapped in."""

    def __init__(self, in_channels=320, out_channels=80, channels=(256,), dropout=0.0, attention_head_dim=64, n_blocks=4,
                 num_mid_blocks=12, num_heads=8, act_fn="gelu", weight_seed=1986, precision="bf16"):
        if num_heads != 8:
            raise NotImplementedError("B200 non-causal estimator: 8 heads x 64 (the fused transformer-block geometry)")
        super().__init__(in_channels=in_channels, out_channels=out_channels, channels=channels, dropout=dropout,
                         attention_head_dim=attention_head_dim, n_blocks=n_blocks, num_mid_blocks=num_mid_blocks,
                         num_heads=num_heads, act_fn=act_fn, weight_seed=weight_seed, precision=precision, _causal=False)

    def run(self, x, mask, mu, t, spks=None, cond=None, streaming=False):
        return super().run(x, mask, mu, t, spks, cond, False)  # the reference class ignores `streaming` (decoder.py:241)


class ConditionalCFM(nn.Module):
    def __init__(self, in_channels, cfm_params, n_spks=1, spk_emb_dim=64, estimator=None):
        super().__init__()
        self.n_feats, self.n_spks, self.spk_emb_dim = in_channels, n_spks, spk_emb_dim
        self.solver = _get(cfm_params, "solver", "euler")
        self.sigma_min = _get(cfm_params, "sigma_min", 1e-4)
        self.t_scheduler = _get(cfm_params, "t_scheduler", "cosine")
        self.training_cfg_rate = _get(cfm_params, "training_cfg_rate", 0.2)
        self.inference_cfg_rate = _get(cfm_params, "inference_cfg_rate", 0.7)
        self.estimator = estimator

    # -- helpers ------------------------------------------------------------------------------
    def _t_span(self, n_timesteps):
        """fp32 tensor of the schedule; ``_t_span_values`` is the same thing as Python floats."""
        return torch.tensor(_t_span_values(int(n_timesteps), self.t_scheduler), dtype=torch.float32)

    def _check_inputs(self, mu, mask, spks, cond, n_timesteps):
        """Shapes are validated here: past this point only raw pointers and sizes cross the C ABI."""
        feat = self.estimator.out_channels
        if mu.dim() != 3 or mu.shape[1] != feat or mu.shape[2] < 1 or n_timesteps < 1:
            raise ValueError(f"mu must be [B, {feat}, T >= 1] and n_timesteps >= 1 (got {tuple(mu.shape)}, "
                             f"{n_timesteps} steps)")
        B, F, T = mu.shape
        if tuple(mask.shape) != (B, 1, T):
            raise ValueError(f"mask must be [{B}, 1, {T}], got {tuple(mask.shape)}")
        if spks is not None and tuple(spks.shape) != (B, F):
            raise ValueError(f"spks must be [{B}, {F}], got {tuple(spks.shape)}")
        if cond is not None and tuple(cond.shape) != (B, F, T):
            raise ValueError(f"cond must be [{B}, {F}, {T}], got {tuple(cond.shape)}")

    def _solve(self, z, t_span, mu, mask, spks, cond, streaming=False):
        """z: [1 or B, 80, >=T] noise rows (row stride may exceed T)."""
        dev = mu.device
        self._check_inputs(mu, mask, spks, cond, len(t_span) - 1)
        B, F, T = mu.shape
        if z.dim() != 3 or z.shape[1] != F or z.shape[2] < T:
            raise ValueError(f"noise must be [1, {F}, >= {T}], got {tuple(z.shape)}")
        _check_prefix_mask(mask, self.estimator.precision == "fp32")
        if z.shape[0] != 1:
            raise NotImplementedError("per-utterance noise: pass z with a single leading row shared by the batch")
        spks = torch.zeros(B, F, device=dev) if spks is None else spks
        cond = torch.zeros_like(mu) if cond is None else cond
        h = self.estimator.handle(dev)
        return torch.ops.ls_b200.flow_solve(h.key, _as_f32(mu, dev), _as_f32(mask, dev), _as_f32(spks, dev),
                                            _as_f32(cond, dev), z[0], [float(v) for v in t_span], 1.0,
                                            float(self.inference_cfg_rate), bool(streaming))

    # -- reference surface --------------------------------------------------------------------
    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, prompt_len=0,
                cache=None, noise=None):
        """flow_matching.py:39-72.  ``noise`` (optional, [1,80,T]) injects the initial sample for parity;
        by default it is drawn with torch.randn like the reference."""
        if mu.shape[0] != 1:
            raise ValueError("the prompt/overlap cache path is defined for batch 1 (flow.py:453)")
        cache = torch.zeros(1, mu.shape[1], 0, 2) if cache is None else cache
        z = (torch.randn_like(mu) if noise is None else noise.to(mu.device, mu.dtype)) * temperature
        z = _as_f32(z, mu.device)
        cache_size = cache.shape[2]
        if cache_size != 0:
            z[:, :, :cache_size] = cache[:, :, :, 0].to(z)
            mu[:, :, :cache_size] = cache[:, :, :, 1].to(mu)  # in place, like the reference (:62-64)
        z_cache = torch.concat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2)
        mu_cache = torch.concat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)
        cache = torch.stack([z_cache, mu_cache.to(z_cache)], dim=-1)
        return self._solve(z, _t_span_values(int(n_timesteps), self.t_scheduler), mu, mask, spks, cond), cache

    def solve_euler(self, x, t_span, mu, mask, spks, cond, streaming=False):
        """flow_matching.py:74-126; ``x`` is the initial noise."""
        return self._solve(_as_f32(x, mu.device), t_span.detach().float().cpu().tolist(), mu, mask, spks, cond, streaming)

    def forward_estimator(self, x, mask, mu, t, spks, cond, streaming=False):
        """flow_matching.py:128-155 (nn.Module branch)."""
        return self.estimator(x, mask, mu, t, spks, cond, streaming=streaming)

    def compute_loss(self, *a, **k):
        raise NotImplementedError("training losses are out of scope of the B200 hot path (SURVEY.md section 8)")

    compute_loss_contrastive = compute_loss


class CausalConditionalCFM(ConditionalCFM):
    def __init__(self, in_channels, cfm_params, n_spks=1, spk_emb_dim=64, estimator=None):
        super().__init__(in_channels, cfm_params, n_spks, spk_emb_dim, estimator)
        # flow_matching.py:320-321 draws seed-0 noise (and reseeds the global RNGs as a side effect, which
        # this implementation deliberately does not do)
        self.rand_noise = synth.fixed_noise()
        self._noise_dev = {}

    def _noise_on(self, device):
        key = str(device)
        if key not in self._noise_dev:
            self._noise_dev[key] = self.rand_noise.to(device=device, dtype=torch.float32).contiguous()
        return self._noise_dev[key]

    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, streaming=False):
        """flow_matching.py:323-348 -> ``(latent [B,80,T] fp32, None)``; B >= 1 (per-utterance semantics)."""
        return self.run(mu, mask, n_timesteps, temperature, spks, cond, streaming)

    def run(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, streaming=False):
        """``forward`` without the inference_mode decorator (the form torch.compile traces)."""
        dev = mu.device
        self._check_inputs(mu, mask, spks, cond, n_timesteps)
        B, F, T = mu.shape
        if T > self.rand_noise.shape[2]:
            raise ValueError(f"T={T} exceeds the fixed-noise buffer ({self.rand_noise.shape[2]} frames)")
        _check_prefix_mask(mask, self.estimator.precision == "fp32")
        spks = torch.zeros(B, F, device=dev) if spks is None else spks
        cond = torch.zeros_like(mu) if cond is None else cond
        h = self.estimator.handle(dev)
        out = torch.ops.ls_b200.flow_solve(h.key, _as_f32(mu, dev), _as_f32(mask, dev), _as_f32(spks, dev),
                                           _as_f32(cond, dev), self._noise_on(dev)[0],
                                           list(_t_span_values(int(n_timesteps), self.t_scheduler)), float(temperature),
                                           float(self.inference_cfg_rate), bool(streaming))
        return out, None
