"""torch custom-op layer over the C ABI (``torch.ops.ls_b200.*``).

north_star asks for host code in Python/PyTorch that "calls hand-written sm_100a CUDA kernels through a thin C-ABI
torch custom-op layer"; SURVEY.md section 8b adds "register a fake/meta kernel so the op composes".  The seam these
ops sit in is the reference's ``ConditionalCFM.forward_estimator`` (speech/cosyvoice/flow/flow_matching.py:128-155:
nn.Module call or TensorRT pointer binding) and ``DACVAE.decode`` (dac-vae/model.py:485-488).

Every op is registered for CUDA only: CPU tensors raise (there is no CPU or PyTorch fallback).  Fake kernels give the
output shapes, so the drop-in modules trace under ``torch.compile(fullgraph=True)`` / ``torch.export``.  Handles (the
packed weights + workspace behind an ``ls_flow*`` / ``ls_dac*``) are passed as integer keys into a registry of live
``native.*Handle`` objects.  Tensor arguments are float32 and contiguous (the drop-in modules coerce before calling);
the ctypes binding in ``native.py`` stays available for hosts without torch op dispatch.
"""
import itertools
import threading
import weakref
from typing import Optional, Sequence

import torch
from torch import Tensor

_handles = weakref.WeakValueDictionary()
_keys = itertools.count(1)
_lock = threading.Lock()


def register_handle(obj) -> int:
    """Give a live native handle object an integer key usable as a custom-op argument."""
    with _lock:
        key = next(_keys)
        _handles[key] = obj
    return key


def _handle(key: int):
    try:
        return _handles[key]
    except KeyError:
        raise RuntimeError(f"ls_b200 handle {key} is not alive (the module that owned it was released or re-loaded)") from None


def _chk(name, t, ndim=None):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 tensor (got {t.dtype}, contiguous={t.is_contiguous()})")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got {tuple(t.shape)}")


# ---- estimator forward: the forward_estimator seam (x, mask, mu, t, spks, cond -> estimator_out) -------------------
@torch.library.custom_op("ls_b200::estimator_forward", mutates_args=(), device_types="cuda")
def estimator_forward(handle: int, x: Tensor, mask: Tensor, mu: Tensor, t: Tensor, spks: Tensor, cond: Tensor,
                      streaming: bool) -> Tensor:
    for n, v in (("x", x), ("mask", mask), ("mu", mu), ("t", t), ("spks", spks), ("cond", cond)):
        _chk(n, v)
    return _handle(handle).estimator_forward(x, mask, mu, t, spks, cond, streaming)


@estimator_forward.register_fake
def _(handle, x, mask, mu, t, spks, cond, streaming):
    return torch.empty_like(x)


# ---- whole Euler/CFG solve (flow_matching.py:74-126) ---------------------------------------------------------------
@torch.library.custom_op("ls_b200::flow_solve", mutates_args=(), device_types="cuda")
def flow_solve(handle: int, mu: Tensor, mask: Tensor, spks: Tensor, cond: Tensor, noise: Tensor,
               t_span: Sequence[float], temperature: float, cfg_rate: float, streaming: bool) -> Tensor:
    """noise: [80, >= T] float32, last dim contiguous (row stride free: a slice of rand_noise)."""
    for n, v in (("mu", mu), ("mask", mask), ("spks", spks), ("cond", cond)):
        _chk(n, v)
    if noise.dtype != torch.float32 or noise.dim() != 2 or noise.stride(-1) != 1:
        raise ValueError("noise must be float32 [80, >= T] with a contiguous last dim")
    return _handle(handle).solve(mu, mask, spks, cond, noise, list(t_span), temperature, cfg_rate, streaming)


@flow_solve.register_fake
def _(handle, mu, mask, spks, cond, noise, t_span, temperature, cfg_rate, streaming):
    return torch.empty_like(mu)


# ---- DAC-VAE decode (dac-vae/model.py:485-488) ---------------------------------------------------------------------
@torch.library.custom_op("ls_b200::dac_decode", mutates_args=(), device_types="cuda")
def dac_decode(handle: int, z: Tensor, lengths: Optional[Tensor], hop: int) -> Tensor:
    _chk("z", z, 3)
    if lengths is not None and (lengths.dtype != torch.int32 or not lengths.is_contiguous()):
        raise ValueError("lengths must be a contiguous int32 tensor")
    h = _handle(handle)
    if hop != h.hop_length:
        raise ValueError(f"hop {hop} does not match the handle's hop length {h.hop_length}")
    return h.decode(z, lengths)


@dac_decode.register_fake
def _(handle, z, lengths, hop):
    return z.new_empty(z.shape[0], 1, z.shape[2] * hop)


# ---- mask [B,1,T] -> valid frames per utterance, int32 [B] (the glue between the solve and the decoder) -------------
@torch.library.custom_op("ls_b200::mask_to_lengths", mutates_args=(), device_types="cuda")
def mask_to_lengths(mask: Tensor) -> Tensor:
    from . import native
    _chk("mask", mask, 3)
    return native.mask_to_lengths(mask)


@mask_to_lengths.register_fake
def _(mask):
    return mask.new_empty(mask.shape[0], dtype=torch.int32)
