"""Drop-in for ``cosyvoice.llm.llm.LearnableSpeakerEncoder`` (speech/cosyvoice/llm/llm.py:34-96; SURVEY.md section 8 row
f-4): reference mel-spectrogram -> L2-normalised speaker embedding, the ``embedding`` input of the flow front half.  Same
constructor and ``forward(x, mask=None)`` signature, same state_dict keys (``init.*``, ``attn.{i}.norm|qkv|proj_out.*``,
``output_proj.*``).  ``precision="bf16"`` (default): tensor cores (conv_gemm + the flash-attention kernel, csrc/front_engine.cu);
``"fp32"``: CUDA-core kernels of csrc/f32_path.cu."""
import torch
import torch.nn as nn

from . import native, synth
from .flow import _as_f32, _register_tree


class LearnableSpeakerEncoder(nn.Module):
    def __init__(self, mel_dim=80, model_dim=512, output_dim=192, num_blocks=6, num_heads=8, dropout=0.0, mean_pooling=False,
                 weight_seed=13, precision="bf16"):
        super().__init__()
        self.precision = native.check_precision(precision)  # "bf16": tensor-core path (csrc/front_engine.cu SpeakerEngine)
        if mean_pooling:
            raise NotImplementedError("first-position pooling only (the reference's default, llm.py:47,88)")
        if model_dim != num_heads * 64 or model_dim % 32:
            raise NotImplementedError("head dim 64 only (llm.py:41-44: 512 / 8)")
        self.mel_dim, self.dim, self.output_dim = mel_dim, model_dim, output_dim
        _register_tree(self, synth.speaker_encoder_state_dict(weight_seed, mel_dim=mel_dim, model_dim=model_dim,
                                                              output_dim=output_dim, num_blocks=num_blocks))
        self._handle = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._handle = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def handle(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("the B200 hot path runs on CUDA tensors only (no CPU fallback)")
        if self._handle is None or self._handle.device != device or self._handle.precision != self.precision:
            self._handle = native.SpeakerHandle(self.state_dict(), device, self.precision)
        return self._handle

    @torch.inference_mode()
    def forward(self, x, mask=None):
        """x: mel [B, 80, T] -> [B, output_dim].  ``mask`` is accepted and ignored, as in the reference (llm.py:70-96: only the
        mean-pooling branch reads it)."""
        if x.dim() != 3 or x.shape[1] != self.mel_dim or x.shape[2] < 1:
            raise ValueError(f"x must be [B, {self.mel_dim}, T >= 1]")
        return self.handle(x.device).encode(_as_f32(x, x.device).unsqueeze(0))

    @torch.inference_mode()
    def encode_references(self, reference_mels):
        """``CausalMaskedDiffWithXvec.get_speaker_embedding`` (flow/flow.py:336-366): [B, 80, T] or [B, N, 80, T] (N reference
        clips, per-clip embeddings averaged) -> L2-normalised [B, output_dim]."""
        if reference_mels.dim() == 3:
            return self.forward(reference_mels)
        if reference_mels.dim() != 4 or reference_mels.shape[2] != self.mel_dim:
            raise ValueError(f"reference_mels must be [B, {self.mel_dim}, T] or [B, N, {self.mel_dim}, T]")
        mel = _as_f32(reference_mels, reference_mels.device).transpose(0, 1).contiguous()  # [N, B, 80, T]
        return self.handle(mel.device).encode(mel)
