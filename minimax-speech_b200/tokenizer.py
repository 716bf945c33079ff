"""FSQ quantizer head of the S3 speech tokenizer (second half).

Drop-in for ``s3tokenizer.model_v2.FSQCodebook`` / ``FSQVectorQuantization``
(speech/tools/S3Tokenizer/s3tokenizer/model_v2.py:83-147): ``encode(hidden [B, T, dim]) -> int32 tokens [B, T]`` with the
reference's parameter names (``project_down.weight [8, dim]``, ``project_down.bias [8]``), so a reference tokenizer
checkpoint's ``quantizer._codebook.*`` entries load unchanged.  The tokens are the 25 Hz FSQ ids (vocabulary 3^8 = 6561)
the flow's ``input_embedding`` consumes.  The whisper-style ``AudioEncoderV2`` trunk that produces ``hidden``
(model_v2.py:243-351) is NOT built: this module starts from its output.
"""
import math

import torch
import torch.nn as nn

from . import native


class FSQCodebook(nn.Module):
    def __init__(self, dim=1280, level=3, weight_seed=0):
        super().__init__()
        if level != 3:
            raise NotImplementedError("level 3 (codebook size 3^8), the reference's only configuration")
        self.level = level
        g = torch.Generator().manual_seed(weight_seed)
        bound = 1.0 / math.sqrt(dim)  # nn.Linear's default initialiser
        self.project_down = nn.Linear(dim, 8)
        with torch.no_grad():
            self.project_down.weight.copy_((torch.rand(8, dim, generator=g) * 2 - 1) * bound)
            self.project_down.bias.copy_((torch.rand(8, generator=g) * 2 - 1) * bound)
        for p in self.parameters():
            p.requires_grad_(False)

    @torch.inference_mode()
    def encode(self, x):
        if x.dim() != 3 or x.shape[-1] != self.project_down.in_features:
            raise ValueError(f"hidden must be [B, T, {self.project_down.in_features}], got {tuple(x.shape)}")
        if x.device.type != "cuda":
            raise RuntimeError("the B200 path runs on CUDA tensors only (no CPU fallback)")
        dev = x.device
        w = self.project_down.weight.to(device=dev, dtype=torch.float32).contiguous()
        b = self.project_down.bias.to(device=dev, dtype=torch.float32).contiguous()
        return native.fsq_encode(x.to(torch.float32).contiguous(), w, b)

    def decode(self, embed_ind):
        raise NotImplementedError("There is no official up project component provided")  # model_v2.py:114-117


class FSQVectorQuantization(nn.Module):
    """model_v2.py:120-147."""

    def __init__(self, dim=1280, codebook_size=3 ** 8, weight_seed=0):
        super().__init__()
        assert 3 ** 8 == codebook_size
        self._codebook = FSQCodebook(dim=dim, level=3, weight_seed=weight_seed)
        self.codebook_size = codebook_size

    def encode(self, x):
        return self._codebook.encode(x)
