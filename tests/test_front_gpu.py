"""Token -> mu front half (SURVEY.md section 8 row f-1, fp32 mode) on the GPU against the reference's golden encoder
outputs and the CPU oracle, then FSQ tokens -> waveform through the whole path."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.dac import DACVAEDecoder  # noqa: E402
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder  # noqa: E402
from minimax_speech_b200.front import TokenToMu  # noqa: E402
from oracle import restatement as O  # noqa: E402

DEV = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)


@pytest.fixture(scope="module")
def front(golden_dir):
    g = np.load(os.path.join(golden_dir, "conformer_golden.npz"))
    sd = synth.conformer_encoder_state_dict(int(g["weights_seed"]))
    f = TokenToMu(precision="fp32")
    f.load_state_dict(sd)
    return g, sd, f


@pytest.fixture(scope="module")
def front_tc(golden_dir):
    """The tensor-core path (bf16 operands, fp32 accumulate): north_star's 1e-2 bar."""
    g = np.load(os.path.join(golden_dir, "conformer_golden.npz"))
    sd = synth.conformer_encoder_state_dict(int(g["weights_seed"]))
    f = TokenToMu()
    assert f.precision == "bf16"
    f.load_state_dict(sd)
    return g, sd, f


def test_mu_vs_reference_golden(front):
    """mu = encoder_proj(h) with h the unmodified reference encoder's output (golden case a, one utterance)."""
    g, sd, f = front
    tok, emb = synth.token_inputs(0, int(g["enc_a_lens"][0]))
    mu, spks = f(tok.to(DEV), emb.to(DEV))
    h = torch.from_numpy(g["enc_a_h"])
    mu_ref = torch.nn.functional.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"]).transpose(1, 2)
    e = O.rel_l2(mu.cpu(), mu_ref)
    print(f"front mu vs reference golden: rel-L2 {e:.3e}")
    assert mu.shape == (1, 80, 80) and e < 1e-4


def test_batch_vs_oracle(front):
    g, sd, f = front
    toks, embs = zip(*[synth.token_inputs(20 + b, 37) for b in range(3)])
    tok, emb = torch.cat(toks, 0), torch.cat(embs, 0)
    mu, spks = f(tok.to(DEV), emb.to(DEV))
    for b in range(3):
        with torch.inference_mode():
            mr, sr = O.tokens_to_mu(sd, tok[b:b + 1], emb[b:b + 1])
        assert O.rel_l2(mu[b:b + 1].cpu(), mr) < 1e-4 and O.rel_l2(spks[b:b + 1].cpu(), sr) < 1e-5


def test_tokens_to_waveform(front):
    """Synthetic FSQ tokens -> mu -> 3-step CFM solve -> DAC decode (fp32 mode throughout) against the oracle."""
    g, sd, f = front
    esd = synth.estimator_state_dict(3, init="test", n_blocks=1, num_mid_blocks=1)
    est = CausalConditionalDecoder(n_blocks=1, num_mid_blocks=1, precision="fp32")
    est.load_state_dict(esd)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    dsd = synth.dac_decoder_state_dict(5, init="test")
    dac = DACVAEDecoder(precision="fp32")
    dac.load_state_dict(dsd)
    tok, emb = synth.token_inputs(9, 25)  # one second of 25 Hz tokens
    lat, _ = f.inference(tok.to(DEV), emb.to(DEV), cfm, n_timesteps=3)
    wav = dac.decode(lat)
    assert lat.shape == (1, 80, 50) and wav.shape == (1, 1, 24000)
    with torch.inference_mode():
        mu, spks = O.tokens_to_mu(sd, tok, emb)
        mask = torch.ones(1, 1, 50)
        lat_ref = O.cfm_forward(esd, synth.fixed_noise(), mu, mask, 3, 1.0, spks, torch.zeros_like(mu))
        wav_ref = O.dac_decode(dsd, lat_ref)
    e, s = O.rel_l2(lat.cpu(), lat_ref), O.snr_db(wav.cpu(), wav_ref)
    print(f"tokens -> waveform (fp32): latent rel-L2 {e:.3e}, waveform SNR {s:.1f} dB")
    assert e < 1e-4 and s > 80.0


def test_non_final_streaming_chunk_vs_reference_golden(front):
    """finalize=False (3 look-ahead context tokens) + streaming=True (block-causal attention) against the unmodified
    reference encoder's output (golden case c) and the oracle."""
    g, sd, f = front
    tok, emb = synth.token_inputs(5, 60)
    mu, spks = f(tok.to(DEV), emb.to(DEV), finalize=False, streaming=True)
    h = torch.from_numpy(g["enc_c_h"])
    mu_ref = torch.nn.functional.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"]).transpose(1, 2)
    e = O.rel_l2(mu.cpu(), mu_ref)
    print(f"front mu (context + streaming) vs reference golden: rel-L2 {e:.3e}")
    assert mu.shape == (1, 80, 114) and e < 1e-4
    with torch.inference_mode():
        mo, _ = O.tokens_to_mu(sd, tok, emb, finalize=False, streaming=True)
    assert O.rel_l2(mu.cpu(), mo) < 1e-4


# ---- tensor-core path ---------------------------------------------------------------------------------------------------
def _mu_ref(g, sd, key):
    h = torch.from_numpy(g[key])
    return torch.nn.functional.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"]).transpose(1, 2)


def test_tc_mu_vs_reference_golden(front_tc):
    g, sd, f = front_tc
    tok, emb = synth.token_inputs(0, int(g["enc_a_lens"][0]))
    mu, spks = f(tok.to(DEV), emb.to(DEV))
    e = O.rel_l2(mu.cpu(), _mu_ref(g, sd, "enc_a_h"))
    print(f"tensor-core front mu vs reference golden: rel-L2 {e:.3e}")
    assert mu.shape == (1, 80, 80) and e < 1e-2
    with torch.inference_mode():
        _, sr = O.tokens_to_mu(sd, tok, emb)
    assert O.rel_l2(spks.cpu(), sr) < 1e-5


def test_tc_non_final_streaming_chunk_vs_reference_golden(front_tc):
    g, sd, f = front_tc
    tok, emb = synth.token_inputs(5, 60)
    mu, _ = f(tok.to(DEV), emb.to(DEV), finalize=False, streaming=True)
    e = O.rel_l2(mu.cpu(), _mu_ref(g, sd, "enc_c_h"))
    print(f"tensor-core front mu (context + streaming) vs reference golden: rel-L2 {e:.3e}")
    assert mu.shape == (1, 80, 114) and e < 1e-2


@pytest.mark.parametrize("n_tokens,batch", [(37, 3), (250, 16), (129, 2)])
def test_tc_batch_vs_oracle_and_batch_invariance(front_tc, n_tokens, batch):
    """Several utterances in one call == each alone (bit-exact), and against the oracle; 250 tokens x 16 = BASELINE configs[1]."""
    g, sd, f = front_tc
    toks, embs = zip(*[synth.token_inputs(20 + b, n_tokens) for b in range(batch)])
    tok, emb = torch.cat(toks, 0), torch.cat(embs, 0)
    mu, spks = f(tok.to(DEV), emb.to(DEV))
    assert mu.shape == (batch, 80, 2 * n_tokens)
    for b in (0, batch - 1):
        with torch.inference_mode():
            mr, sr = O.tokens_to_mu(sd, tok[b:b + 1], emb[b:b + 1])
        e = O.rel_l2(mu[b:b + 1].cpu(), mr)
        print(f"tensor-core front, {n_tokens} tokens x {batch}, utterance {b} vs oracle: rel-L2 {e:.3e}")
        assert e < 1e-2 and O.rel_l2(spks[b:b + 1].cpu(), sr) < 1e-5
    one, _ = f(tok[1:2].to(DEV), emb[1:2].to(DEV))
    assert torch.equal(one, mu[1:2])


def test_tc_streaming_batch_vs_oracle(front_tc):
    g, sd, f = front_tc
    toks, embs = zip(*[synth.token_inputs(70 + b, 103) for b in range(2)])
    tok, emb = torch.cat(toks, 0), torch.cat(embs, 0)
    mu, _ = f(tok.to(DEV), emb.to(DEV), finalize=False, streaming=True)
    with torch.inference_mode():
        mr, _ = O.tokens_to_mu(sd, tok[1:2], emb[1:2], finalize=False, streaming=True)
    e = O.rel_l2(mu[1:2].cpu(), mr)
    print(f"tensor-core front, streaming non-final chunk x 2 vs oracle: rel-L2 {e:.3e}")
    assert mu.shape == (2, 80, 200) and e < 1e-2


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_edge_cases_vs_oracle(golden_dir, precision, tol):
    """One token; a non-final chunk that leaves one token (4 = 1 + 3 context); negative ids clamp to 0 (flow.py:476);
    the all-zero speaker embedding stays zero through the normalise (flow.py:465-466)."""
    sd = synth.conformer_encoder_state_dict(7)
    f = TokenToMu(precision=precision)
    f.load_state_dict(sd)
    cases = []
    tok, emb = synth.token_inputs(90, 1)
    cases.append((tok, emb, True))
    tok, emb = synth.token_inputs(91, 4)
    cases.append((tok, emb, False))
    tok, emb = synth.token_inputs(92, 9)
    tok = tok.clone()
    tok[0, 2], tok[0, 5] = -1, -7
    cases.append((tok, torch.zeros_like(emb), True))
    for tok, emb, fin in cases:
        mu, spks = f(tok.to(DEV), emb.to(DEV), finalize=fin)
        with torch.inference_mode():
            mr, sr = O.tokens_to_mu(sd, tok, emb, finalize=fin)
        assert mu.shape == mr.shape and torch.isfinite(mu).all()
        assert O.rel_l2(mu.cpu(), mr) < tol and O.rel_l2(spks.cpu(), sr) < 1e-5
    with pytest.raises(ValueError):
        f(torch.zeros(1, 3, dtype=torch.int64, device=DEV), emb.to(DEV), finalize=False)  # nothing left after the context


def test_yaml_style_construction():
    """speech/config.yaml:60-116 with the class paths swapped: encoder holder + CFM passed as constructor arguments."""
    from minimax_speech_b200.front import CausalMaskedDiffWithXvec, UpsampleConformerEncoder
    enc = UpsampleConformerEncoder(output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
                                   positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True,
                                   input_layer="linear", pos_enc_layer_type="rel_pos_espnet", selfattention_layer_type="rel_selfattn",
                                   input_size=512, use_cnn_module=False, macaron_style=False, static_chunk_size=25)
    est = CausalConditionalDecoder(**synth.PIPE_EST)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    m = CausalMaskedDiffWithXvec(input_size=512, output_size=80, spk_embed_dim=192, output_type="mel", vocab_size=6561,
                                 input_frame_rate=25, only_mask_loss=True, token_latent_ratio=2, pre_lookahead_len=3,
                                 use_speaker_encoder=False, freeze_speaker_encoder=True, speaker_encoder_path=None, encoder=enc,
                                 decoder=cfm)
    tok, emb = synth.token_inputs(3, 20)
    feat, _ = m.inference(tok.to(DEV), torch.tensor([20]), torch.zeros(1, 0, dtype=torch.int64, device=DEV), torch.tensor([0]),
                          torch.zeros(1, 0, 80, device=DEV), torch.tensor([0]), embedding=emb.to(DEV), finalize=True)
    assert feat.shape == (1, 80, 40) and torch.isfinite(feat).all()
    with pytest.raises(NotImplementedError):
        UpsampleConformerEncoder(input_size=512, output_size=512, macaron_style=True)


@pytest.mark.parametrize("precision,tol", [("bf16", 1e-2), ("fp32", 1e-4)])
def test_padded_batch_vs_reference_golden(golden_dir, precision, tol):
    """A right-padded batch (23 and 15 tokens; golden case b: the unmodified reference encoder run with xs_lens): valid frames
    against the golden output, zeros past 2 * token_len, and the longest utterance equal to its own unpadded run."""
    g = np.load(os.path.join(golden_dir, "conformer_golden.npz"))
    sd = synth.conformer_encoder_state_dict(int(g["weights_seed"]))
    f = TokenToMu(precision=precision)
    f.load_state_dict(sd)
    lens = [int(n) for n in g["enc_b_lens"]]
    toks, embs = zip(*[synth.token_inputs(i, n) for i, n in enumerate(lens)])
    tok = torch.zeros(len(lens), max(lens), dtype=torch.int64)
    for b, t in enumerate(toks):
        tok[b, :t.shape[1]] = t[0]
    tok[1, lens[1]:] = 4321  # whatever sits in the padding must not matter
    emb = torch.cat(embs, 0)
    mu, spks = f(tok.to(DEV), emb.to(DEV), token_len=torch.tensor(lens))
    ref = _mu_ref(g, sd, "enc_b_h")
    for b, n in enumerate(lens):
        e = O.rel_l2(mu[b:b + 1, :, :2 * n].cpu(), ref[b:b + 1, :, :2 * n])
        print(f"padded batch ({precision}), utterance {b} ({n} tokens) vs reference golden: rel-L2 {e:.3e}")
        assert e < tol
        assert float(mu[b, :, 2 * n:].abs().max()) == 0.0 if 2 * n < mu.shape[2] else True
    # (the shorter utterance differs from its own unpadded run, in the reference too: the pre-lookahead convolution reads
    # the embedded padding rows, upsample_encoder.py:66-107; the longest utterance of a batch has no padding)
    one, _ = f(toks[0].to(DEV), embs[0].to(DEV))
    assert O.rel_l2(mu[0:1].cpu(), one.cpu()) < tol * 0.1
    with pytest.raises(ValueError):
        f(tok.to(DEV), emb.to(DEV), token_len=torch.tensor([23, 0]))
