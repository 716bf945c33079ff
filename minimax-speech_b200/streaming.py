"""Streaming synthesis session: the hop scheduling of ``CosyVoice2Model.tts`` (stream branch,
speech/cosyvoice/cli/model.py:336-366) and ``token2wav`` (:285-319) around the drop-in ``CausalMaskedDiffWithXvec.inference``
(``streaming=True``: block-causal attention in the token encoder and in the estimator; non-final calls carry
``pre_lookahead_len`` tokens of look-ahead context), with the DAC-VAE decoder where the reference CLI still calls HiFT
(the CLI is stale with respect to the DAC-VAE latents, SURVEY section 0, so the vocoder side is ours):

* tokens arrive incrementally (``push``); a hop is synthesised as soon as ``token_hop_len + pre_lookahead_len`` new tokens
  are there (the first hop is lengthened by ``prompt_token_pad`` so that hop boundaries fall on the 25-token attention
  chunks, model.py:339-341); every call re-runs the flow on ALL tokens so far and keeps the frames past ``token_offset``
  (model.py:296, 343-352) -- the block-causal masks make the earlier frames independent of the later tokens;
* ``finish`` runs the last call with ``finalize=True`` on whatever remains (model.py:357-366);
* the DAC-VAE decoder is not causal: a chunk is decoded with ``dac_context`` latent frames of left context and the last
  ``dac_context`` frames are held back until the next call has produced their right context (HiFT's mel / source cache and
  cross-fade, model.py:298-311, play this role in the reference).  With ``dac_context`` >= the decoder's reach (15 latent
  frames) the concatenated chunks equal one decode of the whole latent sequence.
"""
import math

import torch


class StreamingSession:
    token_hop_len = 25  # must match the training static_chunk_size (model.py:255-256)

    def __init__(self, flow, dac, prompt_token, prompt_feat, embedding=None, reference_mels=None, dac_context=16):
        """flow: ``front.CausalMaskedDiffWithXvec``; dac: ``dac.DACVAEDecoder``; prompt_token [1,Tp] int, prompt_feat [1,Fp,80]
        (device tensors); embedding [1,192] or reference_mels for the speaker encoder."""
        if prompt_token.dim() != 2 or prompt_token.shape[0] != 1:
            raise ValueError("prompt_token must be [1, Tp]")
        self.flow, self.dac = flow, dac
        self.prompt_token, self.prompt_feat = prompt_token, prompt_feat
        self.embedding, self.reference_mels = embedding, reference_mels
        self.device = prompt_token.device
        self.lookahead = flow.pre_lookahead_len
        self.ratio = flow.token_latent_ratio
        hop, tp = self.token_hop_len, prompt_token.shape[1]
        self.prompt_token_pad = int(math.ceil(tp / hop) * hop - tp)  # model.py:339
        self.dac_context = int(dac_context)
        self.tokens = []
        self.token_offset = 0
        self.latents = torch.zeros(1, flow.output_size, 0, device=self.device)  # every frame synthesised so far
        self.emitted = 0  # latent frames whose audio has been handed out
        self.finished = False

    # -- token2wav (model.py:285-319), flow half -------------------------------------------------------------------------
    def _flow(self, n_tokens, finalize):
        tok = torch.tensor(self.tokens[:n_tokens], dtype=torch.int64, device=self.device).unsqueeze(0)
        n = lambda t: torch.tensor([t.shape[1]], dtype=torch.int32)  # noqa: E731
        lat, _ = self.flow.inference(token=tok, token_len=n(tok), prompt_token=self.prompt_token, prompt_token_len=n(self.prompt_token),
                                     prompt_feat=self.prompt_feat, prompt_feat_len=n(self.prompt_feat), embedding=self.embedding,
                                     reference_mels=self.reference_mels, streaming=True, finalize=finalize)
        new = lat[:, :, self.token_offset * self.ratio:]  # model.py:296
        self.latents = torch.cat([self.latents, new], dim=2)

    # -- vocoder half: DAC-VAE decode with left context and a held-back tail ---------------------------------------------
    def _decode(self, hold):
        total = self.latents.shape[2]
        end = total - hold
        if end <= self.emitted:
            return torch.zeros(1, 0, device=self.device)
        w0 = max(0, self.emitted - self.dac_context)
        wav = self.dac.decode(self.latents[:, :, w0:total].contiguous())
        hop = self.dac.hop_length
        out = wav[:, 0, (self.emitted - w0) * hop:(end - w0) * hop]
        self.emitted = end
        return out

    def _hop_len(self):
        return self.token_hop_len + self.prompt_token_pad if self.token_offset == 0 else self.token_hop_len

    def push(self, tokens):
        """Append newly generated speech tokens; returns the list of waveform chunks [1, n] that became ready."""
        if self.finished:
            raise RuntimeError("session already finished")
        self.tokens.extend(int(t) for t in tokens)
        chunks = []
        while len(self.tokens) - self.token_offset >= self._hop_len() + self.lookahead:  # model.py:343
            this_hop = self._hop_len()
            self._flow(self.token_offset + this_hop + self.lookahead, finalize=False)
            self.token_offset += this_hop
            w = self._decode(hold=self.dac_context)
            if w.shape[1]:
                chunks.append(w)
        return chunks

    def finish(self):
        """The remaining tokens with ``finalize=True`` (model.py:357-366); returns the last waveform chunk."""
        if self.finished:
            raise RuntimeError("session already finished")
        self.finished = True
        if len(self.tokens) > self.token_offset:
            self._flow(len(self.tokens), finalize=True)
            self.token_offset = len(self.tokens)
        return self._decode(hold=0)
