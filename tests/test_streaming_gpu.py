"""Streaming session: the reference CLI's hop scheduling (cli/model.py:336-366) around the
drop-in pipeline.  Properties the block-causal masks guarantee: the frames handed out hop by hop equal the frames of ONE
streaming-mode call on all tokens, and the audio chunks concatenate to one decode of the whole latent sequence."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.dac import DACVAEDecoder  # noqa: E402
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder  # noqa: E402
from minimax_speech_b200.front import CausalMaskedDiffWithXvec  # noqa: E402
from minimax_speech_b200.streaming import StreamingSession  # noqa: E402
from oracle import restatement as O  # noqa: E402

DEV = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)


def build(precision):
    fsd, esd, _ = synth.pipeline_state_dicts()
    est = CausalConditionalDecoder(precision=precision, **synth.PIPE_EST)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    flow = CausalMaskedDiffWithXvec(decoder=cfm, precision=precision)
    full = dict(fsd)
    full.update({"decoder.estimator." + k: v for k, v in esd.items()})
    flow.load_state_dict(full, strict=True)
    dac = DACVAEDecoder(precision=precision)
    dac.load_state_dict(synth.dac_decoder_state_dict(5, init="test"))
    return flow, dac


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-3), ("fp32", 2e-5)])
def test_hops_equal_one_streaming_call(precision, tol):
    flow, dac = build(precision)
    n_tokens = 103
    tok, emb = synth.token_inputs(60, n_tokens)
    ptok, _ = synth.token_inputs(61, 10)
    pfeat = synth.dac_latents(62, 20).transpose(1, 2).contiguous()
    sess = StreamingSession(flow, dac, ptok.to(DEV), pfeat.to(DEV), embedding=emb.to(DEV))
    assert sess.prompt_token_pad == 15
    chunks, calls = [], 0
    for i in range(0, n_tokens, 7):  # tokens trickle in, seven at a time, as from the LLM thread
        got = sess.push(tok[0, i:i + 7].tolist())
        calls += len(got)
        chunks += got
    assert sess.token_offset == 40 + 25 + 25  # hops of 25 (+15 on the first) once 3 look-ahead tokens are there
    chunks.append(sess.finish())
    lat = sess.latents
    assert lat.shape == (1, 80, 2 * n_tokens) and sess.emitted == 2 * n_tokens
    n = lambda t: torch.tensor([t.shape[1]], dtype=torch.int32)  # noqa: E731
    whole, _ = flow.inference(tok.to(DEV), n(tok), ptok.to(DEV), n(ptok), pfeat.to(DEV), n(pfeat), embedding=emb.to(DEV),
                              streaming=True, finalize=True)
    e = O.rel_l2(lat.cpu(), whole.cpu())
    wav = torch.cat(chunks, dim=1)
    ref = dac.decode(lat)[:, 0, :]
    s = O.snr_db(wav.cpu(), ref.cpu())
    print(f"streaming session ({precision}): hop-wise latents vs one streaming call rel-L2 {e:.2e}; chunked audio vs one decode {s:.1f} dB; "
          f"{len(chunks)} chunks")
    assert e < tol
    assert wav.shape == ref.shape and s > 60.0
    with pytest.raises(RuntimeError):
        sess.push([1, 2, 3])


def test_short_utterance_goes_out_in_finish():
    """Fewer tokens than the first hop: nothing is emitted before ``finish`` (model.py:353-366)."""
    flow, dac = build("bf16")
    tok, emb = synth.token_inputs(63, 20)
    ptok, _ = synth.token_inputs(64, 25)
    pfeat = synth.dac_latents(65, 50).transpose(1, 2).contiguous()
    sess = StreamingSession(flow, dac, ptok.to(DEV), pfeat.to(DEV), embedding=emb.to(DEV))
    assert sess.prompt_token_pad == 0 and sess.push(tok[0].tolist()) == []
    wav = sess.finish()
    assert wav.shape == (1, 20 * 2 * 480) and torch.isfinite(wav).all()
