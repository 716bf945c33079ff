"""Pins oracle/restatement.py against outputs of the unmodified reference (tests/golden/*.npz,
made by oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import minimax_speech_b200.synth as synth
from oracle import restatement as O
from oracle.gen_golden import est_inputs

torch.set_num_threads(os.cpu_count() or 1)


@pytest.fixture(scope="module")
def flow(golden_dir):
    g = np.load(os.path.join(golden_dir, "flow_golden.npz"))
    sd = synth.estimator_state_dict(int(g["weights_seed"]), init="test")
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"])), \
        "synthetic weights differ from the ones the goldens were made with"
    return g, sd


@pytest.fixture(scope="module")
def dac(golden_dir):
    g = np.load(os.path.join(golden_dir, "dac_golden.npz"))
    sd = synth.dac_decoder_state_dict(int(g["weights_seed"]), init="test")
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    return g, sd


def test_fixed_noise_matches_reference(flow):
    g, _ = flow
    n = synth.fixed_noise(64)
    assert np.allclose(n[0, :2, :8].numpy(), g["rand_noise_probe"], atol=0, rtol=0)
    assert abs(float(n[0, 0, 0]) - (-1.12584)) < 1e-4  # SURVEY.md §8c determinism anchor


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_estimator_matches_reference(flow, case):
    g, sd = flow
    x, mask, mu, t, spks, cond = est_inputs(list(g[f"est_{case}_lengths"]), int(g[f"est_{case}_seed"]))
    with torch.inference_mode():
        y = O.estimator_forward(sd, x, mask, mu, t, spks, cond, streaming=bool(g[f"est_{case}_streaming"]))
    ref = torch.from_numpy(g[f"est_{case}_y"])
    assert O.rel_l2(y, ref) < 2e-5


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_cfm_solve_matches_reference(flow, case):
    g, sd = flow
    lengths = [int(v) for v in g[f"cfm_{case}_lengths"]]
    mu, mask, spks, cond = synth.batch_inputs(lengths, first_index=50)
    with torch.inference_mode():
        y = O.cfm_forward(sd, synth.fixed_noise(), mu, mask, int(g[f"cfm_{case}_steps"]), 1.0, spks, cond,
                          streaming=bool(g[f"cfm_{case}_streaming"]))
    y = y * mask  # reference per-utterance outputs are zero-padded in the fixture
    ref = torch.from_numpy(g[f"cfm_{case}_y"])
    for b, n in enumerate(lengths):
        assert O.rel_l2(y[b, :, :n], ref[b, :, :n]) < 5e-5
    assert float((y * (1 - mask)).abs().max()) == 0.0


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_dac_decode_matches_reference(dac, case):
    g, sd = dac
    z = synth.dac_latents(int(g[f"dac_{case}_index"]), int(g[f"dac_{case}_frames"]))
    with torch.inference_mode():
        y = O.dac_decode(sd, z)
    ref = torch.from_numpy(g[f"dac_{case}_y"])
    assert y.shape == ref.shape
    assert O.snr_db(y, ref) > 100.0


def test_dac_varlen_is_per_utterance(dac):
    _, sd = dac
    z = torch.cat([synth.dac_latents(5, 6), torch.zeros(1, 80, 6)], 0)
    z[1, :, :3] = synth.dac_latents(6, 3)[0]
    with torch.inference_mode():
        y = O.dac_decode_varlen(sd, z, [6, 3])
        y1 = O.dac_decode(sd, z[1:2, :, :3])
    assert torch.equal(y[1, :, :1440], y1[0])
    assert float(y[1, :, 1440:].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["a", "b"])
def test_dac_encode_matches_reference(golden_dir, case):
    """oracle.dac_encode against DACVAE.encode of the unmodified reference."""
    g = np.load(os.path.join(golden_dir, "dac_enc_golden.npz"))
    sd = synth.dac_encoder_state_dict(int(g["weights_seed"]), init="test")
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    audio = synth.audio_clip(int(g[f"enc_{case}_index"]), int(g[f"enc_{case}_frames"]) * 480)
    with torch.inference_mode():
        z, m, logs = O.dac_encode(sd, audio)
    assert O.rel_l2(m, torch.from_numpy(g[f"enc_{case}_m"])) < 1e-5
    assert O.rel_l2(logs, torch.from_numpy(g[f"enc_{case}_logs"])) < 1e-5
    assert torch.equal(z, m)
    noise = torch.ones_like(m)
    with torch.inference_mode():
        z2, _, _ = O.dac_encode(sd, audio, noise)
    assert torch.allclose(z2, m + torch.exp(logs))


@pytest.mark.parametrize("case", ["a", "b"])
def test_conformer_encoder_matches_reference(golden_dir, case):
    """oracle.upsample_conformer_encode against the unmodified UpsampleConformerEncoder."""
    g = np.load(os.path.join(golden_dir, "conformer_golden.npz"))
    sd = synth.conformer_encoder_state_dict(int(g["weights_seed"]))
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    lens = [int(v) for v in g[f"enc_{case}_lens"]]
    x = torch.zeros(len(lens), max(lens), 512)
    for b, n in enumerate(lens):
        x[b, :n] = torch.nn.functional.embedding(synth.token_inputs(b, n)[0][0], sd["input_embedding.weight"])
    with torch.inference_mode():
        h, l2 = O.upsample_conformer_encode(sd, x, torch.tensor(lens))
    ref = torch.from_numpy(g[f"enc_{case}_h"])
    for b, n in enumerate(lens):
        assert int(l2[b]) == 2 * n
        assert O.rel_l2(h[b, :2 * n], ref[b, :2 * n]) < 1e-5


def test_conformer_encoder_context_and_streaming_matches_reference(golden_dir):
    """Non-final chunk: 3 look-ahead context tokens + block-causal attention (chunk 25 tokens / 50 frames)."""
    g = np.load(os.path.join(golden_dir, "conformer_golden.npz"))
    sd = synth.conformer_encoder_state_dict(int(g["weights_seed"]))
    x = torch.nn.functional.embedding(synth.token_inputs(5, 60)[0], sd["input_embedding.weight"])
    with torch.inference_mode():
        h, l2 = O.upsample_conformer_encode(sd, x[:, :-3], torch.tensor([60]), context=x[:, -3:], streaming=True)
    assert h.shape == (1, 114, 512)
    assert O.rel_l2(h, torch.from_numpy(g["enc_c_h"])) < 1e-5


@pytest.mark.parametrize("case", ["a", "b"])
def test_speaker_encoder_matches_reference(golden_dir, case):
    """oracle.speaker_encode against the unmodified LearnableSpeakerEncoder."""
    g = np.load(os.path.join(golden_dir, "speaker_golden.npz"))
    sd = synth.speaker_encoder_state_dict(int(g["weights_seed"]))
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    mel = torch.cat([synth.reference_mel(i, int(g[f"spk_{case}_frames"])) for i in range(2)], 0)
    with torch.inference_mode():
        y = O.speaker_encode(sd, mel)
    assert O.rel_l2(y, torch.from_numpy(g[f"spk_{case}_y"])) < 1e-5


@pytest.mark.parametrize("case", ["a", "b"])
def test_flow_inference_matches_reference(golden_dir, case):
    """oracle.flow_inference against the unmodified CausalMaskedDiffWithXvec.inference (flow.py:437-511): prompt tokens,
    prompt latents as cond, x-vector (a) or two reference clips through the speaker encoder + non-final streaming chunk (b)."""
    g = np.load(os.path.join(golden_dir, "pipeline_golden.npz"))
    fsd, esd, ssd = synth.pipeline_state_dicts()
    ck = synth.checksum(fsd) + synth.checksum(esd) + synth.checksum(ssd)
    assert abs(ck - float(g["weights_checksum"])) < 1e-6 * abs(ck)
    a = synth.pipeline_inputs(case)
    with torch.inference_mode():
        y = O.flow_inference(fsd, esd, synth.fixed_noise(), a["token"], a["prompt_token"], a["prompt_feat"], embedding=a["embedding"],
                             reference_mels=a["reference_mels"], sd_spk=ssd, streaming=a["streaming"], finalize=a["finalize"])
    ref = torch.from_numpy(g[f"pipe_{case}_y"])
    assert y.shape == ref.shape
    assert O.rel_l2(y, ref) < 1e-5


# ---- round-2 fixtures: the non-causal ConditionalCFM.forward (prompt / overlap cache) and a trained-scale DAC ----
def test_noncausal_cfm_cache_path_matches_reference(golden_dir):
    """flow_matching.py:39-72 on the unmodified reference: first call with an empty cache, second call reusing the
    returned cache (its z and mu frames overwrite the head of the new call's z and mu)."""
    g = np.load(os.path.join(golden_dir, "cfm_nc_golden.npz"))
    sd = synth.estimator_state_dict(int(g["weights_seed"]), init="test")
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    cache = None
    for i in (1, 2):
        T = int(g[f"nc_{i}_T"])
        mu, mask, spks, cond = synth.batch_inputs([T], first_index=int(g[f"nc_{i}_index"]))
        with torch.inference_mode():
            y, cache = O.cfm_forward_cached(sd, torch.from_numpy(g[f"nc_{i}_z"]), mu, mask, int(g["steps"]), 0.8, spks, cond,
                                            prompt_len=int(g["prompt_len"]), cache=cache)
        assert tuple(cache.shape) == (1, 80, int(g["prompt_len"]) + 34, 2)
        assert torch.equal(cache, torch.from_numpy(g[f"nc_{i}_cache"]))
        assert O.rel_l2(y, torch.from_numpy(g[f"nc_{i}_y"])) < 5e-5


@pytest.mark.parametrize("case", ["a", "b"])
def test_dac_trained_scale_matches_reference(golden_dir, case):
    """Snake alpha in [0.5, 2] and activations of O(10) (the regime of a trained checkpoint, layers.py:18-33)."""
    g = np.load(os.path.join(golden_dir, "dac_trained_golden.npz"))
    sd = synth.dac_decoder_state_dict(int(g["weights_seed"]), init="trained")
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    assert float(g["max_abs_alpha_x"]) > 10.0  # the fixture really is in the large-argument regime of sin
    z = synth.dac_latents(int(g[f"dac_{case}_index"]), int(g[f"dac_{case}_frames"]))
    with torch.inference_mode():
        y = O.dac_decode(sd, z)
    assert O.snr_db(y, torch.from_numpy(g[f"dac_{case}_y"])) > 90.0


@pytest.mark.parametrize("case", ["a", "b"])
def test_noncausal_estimator_matches_reference(golden_dir, case):
    """ConditionalDecoder (Conv1d pad 1 + GroupNorm(8) blocks, decoder.py:88-291): the restatement, batched with per-utterance
    GroupNorm statistics, against one reference call per utterance; and the key schema of the drop-in's synthetic weights."""
    import json
    g = np.load(os.path.join(golden_dir, "est_nc_golden.npz"))
    sd = synth.estimator_state_dict(int(g["weights_seed"]), init="test", causal=False)
    assert abs(synth.checksum(sd) - float(g["weights_checksum"])) < 1e-6 * abs(float(g["weights_checksum"]))
    keys = json.load(open(os.path.join(golden_dir, "est_nc_keys.json")))
    assert {k: list(v.shape) for k, v in sd.items()} == keys
    lengths = [int(v) for v in g[f"est_{case}_lengths"]]
    x, mask, mu, t, spks, cond = est_inputs(lengths, int(g[f"est_{case}_seed"]))
    with torch.inference_mode():
        y = O.estimator_forward(sd, x, mask, mu, t, spks, cond)
    ref = torch.from_numpy(g[f"est_{case}_y"])
    for b, n in enumerate(lengths):
        assert O.rel_l2(y[b, :, :n], ref[b, :, :n]) < 2e-5
        assert float(y[b, :, n:].abs().max() if n < y.shape[2] else 0.0) == 0.0


def test_fsq_codebook_matches_reference(golden_dir):
    """FSQCodebook.encode (S3 tokenizer, model_v2.py:83-117): restatement vs the unmodified reference class."""
    from minimax_speech_b200.tokenizer import FSQCodebook
    g = np.load(os.path.join(golden_dir, "fsq_golden.npz"))
    cb = FSQCodebook(dim=1280, weight_seed=int(g["weight_seed"]))
    assert sorted(cb.state_dict().keys()) == [str(k) for k in g["keys"]]
    hidden = torch.randn(3, 50, 1280, generator=torch.Generator().manual_seed(int(g["hidden_seed"]))) * 3.0
    tok = O.fsq_encode(cb.project_down.weight, cb.project_down.bias, hidden)
    assert torch.equal(tok, torch.from_numpy(g["tokens"]))
    assert int(tok.min()) >= 0 and int(tok.max()) < 3 ** 8


def test_s3_tokenizer_matches_reference(golden_dir):
    """S3TokenizerV2 (encoder trunk + FSQ head, model_v2.py:290-415): restatement vs the unmodified reference class on a
    ragged batch; the drop-in module's parameter names are the reference's."""
    from minimax_speech_b200.tokenizer import S3TokenizerV2

    g = np.load(os.path.join(golden_dir, "s3_golden.npz"))
    n_mels, n_state, n_head, n_layer = [int(v) for v in g["cfg"]]
    sd = synth.s3_tokenizer_state_dict(int(g["weights_seed"]), n_mels, n_state, n_head, n_layer)
    assert sorted(sd.keys()) == [str(k) for k in g["keys"]]

    class Cfg:
        n_audio_state, n_audio_head, n_audio_layer = n_state, n_head, n_layer
    Cfg.n_mels = n_mels
    mod = S3TokenizerV2("speech_tokenizer_v2_25hz", Cfg())
    mod.load_state_dict(sd, strict=True)
    lens = [int(v) for v in g["mel_len"]]
    mel = torch.cat([synth.s3_mel(i, int(g["frames"])) for i in range(len(lens))], 0)
    with torch.inference_mode():
        hidden, code_len = O.s3_encode(sd, mel, torch.tensor(lens))
        codes, _ = O.s3_quantize(sd, mel, torch.tensor(lens))
    assert code_len.tolist() == g["code_len"].tolist()
    ref_h, ref_c = torch.from_numpy(g["hidden"]), torch.from_numpy(g["codes"])
    for b, n in enumerate(code_len.tolist()):
        assert O.rel_l2(hidden[b, :n], ref_h[b, :n]) < 2e-6
        assert torch.equal(codes[b, :n], ref_c[b, :n])


def test_s3_tokenizer_long_clips_match_reference(golden_dir):
    """Batches with clips longer than 30 s: windowing + merging (model_v2.py:417-588, utils.py:367-390) restated."""
    g = np.load(os.path.join(golden_dir, "s3_long_golden.npz"))
    n_mels, n_state, n_head, n_layer = [int(v) for v in g["cfg"]]
    sd = synth.s3_tokenizer_state_dict(int(g["weights_seed"]), n_mels, n_state, n_head, n_layer)
    lens = [int(v) for v in g["mel_len"]]
    mel = torch.zeros(len(lens), n_mels, max(lens))
    for i, n in enumerate(lens):
        mel[i, :, :n] = synth.s3_mel(40 + i, n)[0]
    with torch.inference_mode():
        codes, code_len = O.s3_quantize(sd, mel, torch.tensor(lens))
    assert code_len.tolist() == g["code_len"].tolist() and codes.dtype == torch.long
    assert torch.equal(codes, torch.from_numpy(g["codes"]))
