"""Token -> mu front half of ``CausalMaskedDiffWithXvec.inference`` (speech/cosyvoice/flow/flow.py:437-511; SURVEY.md
section 8 row f-1): speaker-embedding normalise + ``spk_embed_affine_layer``, ``input_embedding``, the
``UpsampleConformerEncoder`` (transformer/upsample_encoder.py) and ``encoder_proj``.  Parameters are registered under the
reference's state_dict keys, so the flow checkpoint loads unchanged.  fp32 mode only in this round (CUDA-core kernels of
csrc/f32_path.cu): equal-length batches, no prompt; final chunks (finalize=True) and non-final chunks (3 look-ahead
context tokens), optional block-causal streaming attention."""
import torch
import torch.nn as nn

from . import native, synth
from .flow import _as_f32, _register_tree


class TokenToMu(nn.Module):
    def __init__(self, input_size=512, output_size=80, spk_embed_dim=192, vocab_size=6561, attention_heads=8,
                 linear_units=2048, num_blocks=6, weight_seed=7, precision="fp32", **_ignored):
        super().__init__()
        if precision != "fp32":
            raise NotImplementedError("the token -> mu front half runs in fp32 mode only (tensor-core path: not built yet)")
        if input_size != attention_heads * 64:
            raise NotImplementedError("head dim 64 only (config.yaml:73-88: 512 / 8)")
        self.precision, self.output_size, self.spk_embed_dim, self.vocab_size = precision, output_size, spk_embed_dim, vocab_size
        _register_tree(self, synth.conformer_encoder_state_dict(weight_seed, d=input_size, heads=attention_heads, ff=linear_units,
                                                                num_blocks=num_blocks, out_dim=output_size, vocab=vocab_size,
                                                                spk_dim=spk_embed_dim))
        self._handle = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._handle = None
        keep = {k: v for k, v in state_dict.items()
                if k.startswith(("input_embedding.", "encoder.", "encoder_proj.", "spk_embed_affine_layer."))}
        return super().load_state_dict(keep, strict=strict, **kw)

    def handle(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the B200 hot path runs on CUDA tensors only (no CPU fallback)")
        if self._handle is None or self._handle.device != device:
            self._handle = native.FrontHandle(self.state_dict(), device)
        return self._handle

    pre_lookahead_len = 3

    @torch.inference_mode()
    def forward(self, token, embedding, finalize=True, streaming=False):
        """token [B,T] int64, embedding [B,192] -> (mu [B,80,2T'], spks [B,80]): the ``mu`` / ``spks`` the reference hands to
        ``self.decoder`` (flow.py:501-508).  finalize=False: the last 3 tokens are look-ahead context (T' = T - 3)."""
        n_ctx = 0 if finalize else self.pre_lookahead_len
        if token.dim() != 2 or token.shape[1] < 1 + n_ctx or token.dtype != torch.int64:
            raise ValueError(f"token must be an int64 tensor [B, T >= {1 + n_ctx}]")
        if tuple(embedding.shape) != (token.shape[0], self.spk_embed_dim):
            raise ValueError(f"embedding must be [{token.shape[0]}, {self.spk_embed_dim}]")
        dev = token.device
        return self.handle(dev).encode(token.contiguous(), _as_f32(embedding, dev), n_ctx, streaming)

    @torch.inference_mode()
    def inference(self, token, embedding, decoder, n_timesteps=10, streaming=False, finalize=True):
        """``CausalMaskedDiffWithXvec.inference`` without prompt (flow.py:437-511): tokens -> latents [B,80,2T']."""
        mu, spks = self.forward(token, embedding, finalize=finalize, streaming=streaming)
        mask = torch.ones(mu.shape[0], 1, mu.shape[2], device=mu.device)
        feat, _ = decoder(mu=mu, mask=mask, spks=spks, cond=torch.zeros_like(mu), n_timesteps=n_timesteps, streaming=streaming)
        return feat.float(), None
