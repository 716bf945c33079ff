"""Drop-in DAC-VAE decoder: ``decode(z [B,latent,L]) -> [B,1,L*hop]`` (dac-vae/model.py:485-488).

``DACVAEDecoder`` carries the ``decoder.*`` / ``de_conv_pre.*`` half of the reference ``DACVAE``
state_dict (weight_g / weight_v / bias / alpha keys) and loads reference
checkpoints (``ckpt['generator']``, dac-vae/inference.py:42-46) with ``strict=False`` semantics for the
encoder keys it does not own.  ``patch_reference_model`` swaps ``decode`` on an instance of the reference
class instead.  The encoder (``encode``) is out of scope.
"""
import torch
import torch.nn as nn

from . import native, synth
from .flow import _as_f32, _register_tree


class DACVAEDecoder(nn.Module):
    def __init__(self, latent_dim=80, decoder_dim=1536, decoder_rates=(5, 4, 4, 3, 2), sample_rate=24000,
                 d_out=1, weight_seed=0, precision="bf16", **_ignored):
        super().__init__()
        self.precision = native.check_precision(precision)
        if d_out != 1:
            raise NotImplementedError("mono output only (configx2.yml: d_out=1)")
        self.latent_dim, self.decoder_dim, self.decoder_rates = latent_dim, decoder_dim, list(decoder_rates)
        self.sample_rate = sample_rate
        self.hop_length = 1
        for r in decoder_rates:
            self.hop_length *= r
        _register_tree(self, synth.dac_decoder_state_dict(weight_seed, "reference", latent_dim=latent_dim,
                                                          decoder_dim=decoder_dim, decoder_rates=decoder_rates))
        self._handle = None
        self._register_load_state_dict_pre_hook(lambda *a, **k: self.invalidate())

    @property
    def device(self):
        return next(self.parameters()).device

    def invalidate(self):
        self._handle = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        own = {k: v for k, v in state_dict.items() if k.startswith(("decoder.", "de_conv_pre."))}
        return super().load_state_dict(own, strict=strict, **kw)

    def handle(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the B200 hot path runs on CUDA tensors only (no CPU fallback)")
        if self._handle is None or self._handle.device != device or self._handle.precision != self.precision:
            self._handle = native.DacHandle(self.state_dict(), device, self.precision)
        return self._handle

    @torch.inference_mode()
    def decode(self, z, lengths=None):
        """``lengths`` (optional int tensor [B]): valid latent frames per item of a right-padded batch; each item
        is then decoded exactly as if alone (the reference decodes one utterance per call)."""
        dev = z.device
        if z.dim() != 3 or z.shape[1] != self.latent_dim or z.shape[2] < 1:
            raise ValueError(f"z must be [B, {self.latent_dim}, L >= 1], got {tuple(z.shape)}")
        if lengths is not None:
            if lengths.numel() != z.shape[0]:
                raise ValueError(f"lengths must have {z.shape[0]} entries")
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        return self.handle(dev).decode(_as_f32(z, dev), lengths)

    forward = decode


def patch_reference_model(model):
    """Replace ``decode`` of a reference ``DACVAE`` instance by the B200 path (weights taken from it)."""
    sd = {k: v for k, v in model.state_dict().items() if k.startswith(("decoder.", "de_conv_pre."))}
    dec = DACVAEDecoder(latent_dim=model.latent_dim, decoder_dim=model.decoder_dim,
                        decoder_rates=tuple(model.decoder_rates), sample_rate=model.sample_rate)
    dec.load_state_dict(sd)
    model.decode = dec.decode
    model._b200_decoder = dec
    return model
