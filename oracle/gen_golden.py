"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_import.py) with the deterministic synthetic
weights of minimax-speech_b200/synth.py loaded into the reference modules.

Run in the build container only:   python -m oracle.gen_golden
The fixtures hold inputs' seeds + reference outputs (small); weights are regenerated from
their seed wherever the fixtures are consumed, guarded by a checksum.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import as R  # noqa: E402
import minimax_speech_b200.synth as synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
EST_SEED, DAC_SEED = 7, 11
ENC_SEED = 7
SPK_SEED = 13


def est_inputs(lengths, seed):
    g = torch.Generator().manual_seed(seed)
    B, T = len(lengths), max(lengths)
    x = torch.randn(B, 80, T, generator=g)
    mu = torch.randn(B, 80, T, generator=g)
    cond = torch.randn(B, 80, T, generator=g)
    spks = torch.randn(B, 80, generator=g)
    t = torch.rand(B, generator=g)
    mask = torch.zeros(B, 1, T)
    for b, n in enumerate(lengths):
        mask[b, :, :n] = 1
    return x, mask, mu, t, spks, cond


def pipeline_golden():
    """The unmodified CausalMaskedDiffWithXvec.inference (flow.py:437-511): prompt tokens + prompt latents + speaker encoder."""
    with torch.inference_mode():
        m = R.build_reference_pipeline(synth.PIPE_EST)
        fsd, esd, ssd = synth.pipeline_state_dicts()
        full = dict(fsd)
        full.update({"decoder.estimator." + k: v for k, v in esd.items()})
        full.update({"speaker_encoder." + k: v for k, v in ssd.items()})
        missing, unexpected = m.load_state_dict(full, strict=False)
        assert not unexpected and not missing, (missing[:5], unexpected[:5])
        out = {"weights_checksum": synth.checksum(fsd) + synth.checksum(esd) + synth.checksum(ssd)}
        import json
        with open(os.path.join(OUT, "pipeline_keys.json"), "w") as f:  # the reference pipeline's state_dict schema
            json.dump({k: list(v.shape) for k, v in m.state_dict().items()}, f, indent=0, sort_keys=True)
        for case in ("a", "b"):
            a = synth.pipeline_inputs(case)
            n = lambda t: torch.tensor([t.shape[1]], dtype=torch.int32)
            kw = {}
            if a["reference_mels"] is not None:
                kw = dict(reference_mels=a["reference_mels"], reference_mel_masks=torch.ones(1, 2, a["reference_mels"].shape[-1]))
            feat, _ = m.inference(a["token"], n(a["token"]), a["prompt_token"], n(a["prompt_token"]), a["prompt_feat"],
                                  n(a["prompt_feat"]), embedding=a["embedding"], streaming=a["streaming"], finalize=a["finalize"],
                                  **kw)
            out[f"pipe_{case}_y"] = feat.numpy()
            print("pipeline", case, feat.shape, float(feat.abs().mean()))
        np.savez_compressed(os.path.join(OUT, "pipeline_golden.npz"), **out)


NC_SEED_1, NC_SEED_2 = 501, 502   # torch.manual_seed before each non-causal forward (its z = torch.randn_like(mu))
DAC_TRAINED_SEED = 21
S3_LONG_LENS = [5700, 900, 3100]  # s3_long_golden.npz: 57 s (3 windows), 9 s, 31 s (2 windows)
S3_SEED, S3_FRAMES, S3_LENS = 21, 203, [203, 150, 96]  # s3_golden.npz: 100 Hz mel frames per utterance (right-padded batch)


def extra_golden():
    """Round-2 fixtures (own files: the round-1 fixtures stay byte-identical).

    cfm_nc_golden.npz     the non-causal twin ``ConditionalCFM.forward`` (flow_matching.py:39-72) called twice (n_timesteps = 10, the value flow.py:192,506 pass) on the
                          unmodified reference: prompt_len = 20 with an empty cache, then with the returned cache
                          (54 frames of z | mu) reused on a longer utterance -- the CLI's streaming overlap path.
    est_nc_golden.npz     the non-causal ConditionalDecoder estimator (decoder.py:88-291), one call per utterance.
    s3_golden.npz         S3TokenizerV2 encoder trunk + quantize (model_v2.py:290-415), reduced width, ragged batch of 3.
    s3_long_golden.npz    the same tokenizer on a batch with clips longer than 30 s (sliding windows, model_v2.py:417-588).
    fsq_golden.npz        FSQCodebook.encode of the S3 tokenizer (tools/S3Tokenizer/s3tokenizer/model_v2.py:83-117).
    dac_trained_golden.npz  DACVAE.decode with weights in the regime of a TRAINED checkpoint (synth init="trained":
                          Snake alpha in [0.5, 2], activations of O(10), |alpha * x| up to ~25 rad), layers.py:18-33.
    """
    os.makedirs(OUT, exist_ok=True)
    with torch.inference_mode():
        sd = synth.estimator_state_dict(EST_SEED, init="test")
        cfm = R.build_reference_flow()
        cfm.estimator.load_state_dict(sd, strict=True)
        nc_forward = type(cfm).__mro__[1].forward  # ConditionalCFM.forward, the parent's (non-causal) method
        assert type(cfm).__mro__[1].__name__ == "ConditionalCFM"
        out = {"weights_seed": EST_SEED, "weights_checksum": synth.checksum(sd), "prompt_len": 20, "steps": 10}
        cache = torch.zeros(1, 80, 0, 2)
        for i, (T, seed) in enumerate([(70, NC_SEED_1), (90, NC_SEED_2)], 1):
            mu, mask, spks, cond = synth.batch_inputs([T], first_index=70 + i)
            torch.manual_seed(seed)
            z = torch.randn_like(mu)  # what the forward below draws (same seed, same shape)
            torch.manual_seed(seed)
            y, cache = nc_forward(cfm, mu.clone(), mask, 10, temperature=0.8, spks=spks, cond=cond, prompt_len=20, cache=cache)
            out[f"nc_{i}_T"], out[f"nc_{i}_index"] = T, 70 + i
            out[f"nc_{i}_z"], out[f"nc_{i}_y"], out[f"nc_{i}_cache"] = z.numpy(), y.numpy(), cache.numpy()
            print("cfm non-causal", i, y.shape, cache.shape, float(y.abs().mean()))
        np.savez_compressed(os.path.join(OUT, "cfm_nc_golden.npz"), **out)

        # ---- the non-causal ConditionalDecoder (decoder.py:88-291): Conv1d(pad 1) + GroupNorm(8) blocks.  One reference
        # call per utterance at its own length (the reference's solve_euler is batch 1): GroupNorm statistics run over the
        # frames of that utterance only.
        sd = synth.estimator_state_dict(EST_SEED + 2, init="test", causal=False)
        est = R.build_reference_noncausal_estimator()
        est.load_state_dict(sd, strict=True)
        out = {"weights_seed": EST_SEED + 2, "weights_checksum": synth.checksum(sd)}
        for name, lengths, seed in [("a", [96], 110), ("b", [130, 77], 111)]:
            x, mask, mu, t, spks, cond = est_inputs(lengths, seed)
            ys = []
            for b, n in enumerate(lengths):
                y = est(x[b:b + 1, :, :n], mask[b:b + 1, :, :n], mu[b:b + 1, :, :n], t[b:b + 1], spks[b:b + 1], cond[b:b + 1, :, :n])
                yp = torch.zeros(1, 80, max(lengths))
                yp[:, :, :n] = y
                ys.append(yp)
            out[f"est_{name}_lengths"], out[f"est_{name}_seed"] = np.array(lengths), seed
            out[f"est_{name}_y"] = torch.cat(ys, 0).numpy()
            print("non-causal estimator", name, out[f"est_{name}_y"].shape, float(np.abs(out[f"est_{name}_y"]).mean()))
        np.savez_compressed(os.path.join(OUT, "est_nc_golden.npz"), **out)
        with open(os.path.join(OUT, "est_nc_keys.json"), "w") as f:
            import json
            json.dump({k: list(v.shape) for k, v in est.state_dict().items()}, f, indent=0, sort_keys=True)

        # ---- FSQ quantizer head of the S3 tokenizer (model_v2.py:83-117) on synthetic hidden states
        from minimax_speech_b200.tokenizer import FSQCodebook as OurFSQ
        ours = OurFSQ(dim=1280, weight_seed=5)
        ref_cb = R.load_fsq_codebook_class()(dim=1280, level=3)
        ref_cb.load_state_dict(ours.state_dict(), strict=True)
        hidden = torch.randn(3, 50, 1280, generator=torch.Generator().manual_seed(17)) * 3.0
        np.savez_compressed(os.path.join(OUT, "fsq_golden.npz"), tokens=ref_cb.encode(hidden).numpy(), hidden_seed=17,
                            weight_seed=5, keys=np.array(sorted(ref_cb.state_dict().keys())))
        print("fsq", ref_cb.encode(hidden)[0, :8].tolist())

        # ---- S3TokenizerV2 (model_v2.py:290-415): encoder trunk + FSQ head, reduced width (2 heads of 64, 2 layers), ragged batch
        s3cfg = dict(n_mels=128, n_state=128, n_head=2, n_layer=2)
        s3sd = synth.s3_tokenizer_state_dict(S3_SEED, **s3cfg)
        s3 = R.build_reference_s3_tokenizer(**s3cfg)
        s3.load_state_dict(s3sd, strict=True)
        mel = torch.cat([synth.s3_mel(i, S3_FRAMES) for i in range(len(S3_LENS))], 0)
        mel_len = torch.tensor(S3_LENS)
        hidden, code_len = s3.encoder(mel, mel_len)
        codes, code_len2 = s3.quantize(mel, mel_len)
        assert torch.equal(code_len, code_len2)
        np.savez_compressed(os.path.join(OUT, "s3_golden.npz"), hidden=hidden.numpy(), codes=codes.numpy(), code_len=code_len.numpy(),
                            mel_len=np.array(S3_LENS), frames=S3_FRAMES, weights_seed=S3_SEED,
                            cfg=np.array([s3cfg[k] for k in ("n_mels", "n_state", "n_head", "n_layer")]),
                            keys=np.array(sorted(s3.state_dict().keys())))
        print("s3 tokenizer", tuple(hidden.shape), code_len.tolist(), codes[1, :6].tolist())
        # clips longer than 30 s: the sliding-window path (_quantize_mixed_batch, model_v2.py:417-588), mixed with a short one
        long_mel = torch.zeros(len(S3_LONG_LENS), s3cfg["n_mels"], max(S3_LONG_LENS))
        for i, n in enumerate(S3_LONG_LENS):
            long_mel[i, :, :n] = synth.s3_mel(40 + i, n)[0]
        lcodes, llen = s3.quantize(long_mel, torch.tensor(S3_LONG_LENS))
        np.savez_compressed(os.path.join(OUT, "s3_long_golden.npz"), codes=lcodes.numpy(), code_len=llen.numpy(),
                            mel_len=np.array(S3_LONG_LENS), weights_seed=S3_SEED,
                            cfg=np.array([s3cfg[k] for k in ("n_mels", "n_state", "n_head", "n_layer")]))
        print("s3 tokenizer, long clips", tuple(lcodes.shape), llen.tolist())

        sd = synth.dac_decoder_state_dict(DAC_TRAINED_SEED, init="trained")
        dac = R.build_reference_dac()
        missing, unexpected = dac.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(("encoder.", "en_conv_post.")) for k in missing), (missing[:5], unexpected)
        out = {"weights_seed": DAC_TRAINED_SEED, "weights_checksum": synth.checksum(sd)}
        peak = [0.0]
        for n, m in dac.named_modules():
            if m.__class__.__name__ == "Snake1d" and n.startswith("decoder"):
                m.register_forward_hook(lambda mod, inp, o: peak.__setitem__(0, max(peak[0], float((inp[0].abs() * mod.alpha.abs()).max()))))
        for name, frames, idx in [("a", 20, 3), ("b", 5, 4)]:
            z = synth.dac_latents(idx, frames)
            y = dac.decode(z)
            out[f"dac_{name}_frames"], out[f"dac_{name}_index"], out[f"dac_{name}_y"] = frames, idx, y.numpy()
            print("dac trained-scale", name, y.shape, float(y.abs().max()), float(y.abs().mean()))
        out["max_abs_alpha_x"] = peak[0]
        print("max |alpha * x| at a Snake input:", peak[0])
        np.savez_compressed(os.path.join(OUT, "dac_trained_golden.npz"), **out)


def main():
    if "--pipeline-only" in sys.argv:
        return pipeline_golden()
    if "--extra-only" in sys.argv:
        return extra_golden()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    with torch.inference_mode():
        # ---- estimator / CFM ----
        sd = synth.estimator_state_dict(EST_SEED, init="test")
        cfm = R.build_reference_flow()
        cfm.estimator.load_state_dict(sd, strict=True)
        est = cfm.estimator
        out = {"weights_seed": EST_SEED, "weights_checksum": synth.checksum(sd)}
        for name, lengths, seed, streaming in [("a", [64, 64], 100, False), ("b", [130, 77], 101, False),
                                               ("c", [130, 77], 102, True)]:
            x, mask, mu, t, spks, cond = est_inputs(lengths, seed)
            y = est(x, mask, mu, t, spks, cond, streaming=streaming)
            out[f"est_{name}_lengths"] = np.array(lengths)
            out[f"est_{name}_seed"] = seed
            out[f"est_{name}_streaming"] = streaming
            out[f"est_{name}_y"] = y.numpy()
            print("estimator", name, y.shape, float(y.abs().mean()))
        # whole solve, literal reference call (B=1 each: flow_matching.py:97-110 is batch-1 only)
        for name, lengths, n_steps, streaming in [("a", [96], 4, False), ("b", [80, 50], 3, False),
                                                  ("c", [120], 2, True)]:
            mu, mask, spks, cond = synth.batch_inputs(lengths, first_index=50)
            ys = []
            for b, n in enumerate(lengths):
                y, _ = cfm(mu=mu[b:b + 1, :, :n].clone(), mask=mask[b:b + 1, :, :n], n_timesteps=n_steps,
                           temperature=1.0, spks=spks[b:b + 1], cond=cond[b:b + 1, :, :n], streaming=streaming)
                yp = torch.zeros(1, 80, max(lengths))
                yp[:, :, :n] = y
                ys.append(yp)
            y = torch.cat(ys, 0)
            out[f"cfm_{name}_lengths"] = np.array(lengths)
            out[f"cfm_{name}_steps"] = n_steps
            out[f"cfm_{name}_streaming"] = streaming
            out[f"cfm_{name}_y"] = y.numpy()
            print("cfm", name, y.shape, float(y.abs().mean()))
        out["rand_noise_probe"] = cfm.rand_noise[0, :2, :8].numpy()
        np.savez_compressed(os.path.join(OUT, "flow_golden.npz"), **out)

        # ---- DAC-VAE decode ----
        sd = synth.dac_decoder_state_dict(DAC_SEED, init="test")
        dac = R.build_reference_dac()
        missing, unexpected = dac.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.startswith(("encoder.", "en_conv_post.")) for k in missing), (missing[:5], unexpected)
        out = {"weights_seed": DAC_SEED, "weights_checksum": synth.checksum(sd)}
        for name, frames, idx in [("a", 24, 0), ("b", 7, 1), ("c", 1, 2)]:
            z = synth.dac_latents(idx, frames)
            y = dac.decode(z)
            out[f"dac_{name}_frames"] = frames
            out[f"dac_{name}_index"] = idx
            out[f"dac_{name}_y"] = y.numpy()
            print("dac", name, y.shape, float(y.abs().max()))
        np.savez_compressed(os.path.join(OUT, "dac_golden.npz"), **out)

        # ---- DAC-VAE encode (model.py:469-483): m and logs are deterministic; z = m + randn * exp(logs) is not stored
        esd = synth.dac_encoder_state_dict(DAC_SEED + 1, init="test")
        missing, unexpected = dac.load_state_dict(esd, strict=False)
        assert not unexpected and all(k.startswith(("decoder.", "de_conv_pre.")) for k in missing), (missing[:5], unexpected)
        out = {"weights_seed": DAC_SEED + 1, "weights_checksum": synth.checksum(esd)}
        import contextlib, io
        for name, frames, idx in [("a", 12, 0), ("b", 3, 1)]:
            audio = synth.audio_clip(idx, frames * 480)
            with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a shape
                z, m, logs = dac.encode(audio)
            out[f"enc_{name}_frames"] = frames
            out[f"enc_{name}_index"] = idx
            out[f"enc_{name}_m"] = m.numpy()
            out[f"enc_{name}_logs"] = logs.numpy()
            print("enc", name, m.shape, float(m.abs().mean()), float(logs.abs().mean()))
        np.savez_compressed(os.path.join(OUT, "dac_enc_golden.npz"), **out)

        # ---- token -> mu front half: the reference UpsampleConformerEncoder on embedded tokens
        import importlib
        R.install_stubs()
        um = importlib.import_module("cosyvoice.transformer.upsample_encoder")
        enc = um.UpsampleConformerEncoder(
            input_size=512, output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
            positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True, input_layer="linear",
            pos_enc_layer_type="rel_pos_espnet", selfattention_layer_type="rel_selfattn", use_cnn_module=False,
            macaron_style=False, static_chunk_size=25).eval()  # config.yaml:73-88
        csd = synth.conformer_encoder_state_dict(ENC_SEED)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in csd.items() if k.startswith("encoder.")}, strict=True)
        out = {"weights_seed": ENC_SEED, "weights_checksum": synth.checksum(csd)}
        for name, lens in [("a", [40]), ("b", [23, 15])]:
            toks = [synth.token_inputs(i, n)[0] for i, n in enumerate(lens)]
            x = torch.zeros(len(lens), max(lens), 512)
            for b, t in enumerate(toks):
                x[b, :t.shape[1]] = torch.nn.functional.embedding(t[0], csd["input_embedding.weight"])
            h, _ = enc(x, torch.tensor(lens))
            out[f"enc_{name}_lens"] = np.array(lens)
            out[f"enc_{name}_h"] = h.numpy()
            print("conformer", name, h.shape, float(h.abs().mean()))
        # non-final streaming chunk: look-ahead context = last 3 tokens, block-causal attention (flow.py:482-489)
        tok = synth.token_inputs(5, 60)[0]
        x = torch.nn.functional.embedding(tok, csd["input_embedding.weight"])
        h, _ = enc(x[:, :-3], torch.tensor([60]), context=x[:, -3:], streaming=True)
        out["enc_c_h"] = h.numpy()
        print("conformer c (context + streaming)", h.shape, float(h.abs().mean()))
        np.savez_compressed(os.path.join(OUT, "conformer_golden.npz"), **out)

        # ---- speaker encoder: the unmodified LearnableSpeakerEncoder
        spk = R.build_reference_speaker_encoder()
        ssd = synth.speaker_encoder_state_dict(SPK_SEED)
        spk.load_state_dict(ssd, strict=True)
        out = {"weights_seed": SPK_SEED, "weights_checksum": synth.checksum(ssd)}
        for name, frames in [("a", 150), ("b", 37)]:
            mel = torch.cat([synth.reference_mel(i, frames) for i in range(2)], 0)
            out[f"spk_{name}_frames"] = frames
            out[f"spk_{name}_y"] = spk(mel).numpy()
            print("speaker", name, out[f"spk_{name}_y"].shape)
        np.savez_compressed(os.path.join(OUT, "speaker_golden.npz"), **out)

        pipeline_golden()
        extra_golden()

        # ---- key schema of the reference state_dicts (drop-in modules must expose exactly these) ----
        import json
        keys = {"estimator": {k: list(v.shape) for k, v in est.state_dict().items()},
                "dac_decoder": {k: list(v.shape) for k, v in dac.state_dict().items()
                                if k.startswith(("decoder.", "de_conv_pre."))},
                "dac_encoder": {k: list(v.shape) for k, v in dac.state_dict().items()
                                if k.startswith(("encoder.", "en_conv_post."))}}
        with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
            json.dump(keys, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
