"""DAC-VAE encoder on the GPU, tensor-core path (bf16 operands: rel-L2 <= 1e-2) and fp32
mode (<= 1e-4), against the reference's golden outputs, the CPU oracle, and through the encode -> decode round trip of
the drop-in modules."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.dac import DACVAEDecoder, DACVAEEncoder  # noqa: E402
from oracle import restatement as O  # noqa: E402

DEV = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)


TOL = {"bf16": 1e-2, "fp32": 1e-4}


@pytest.fixture(scope="module", params=["bf16", "fp32"])
def enc(golden_dir, request):
    g = np.load(os.path.join(golden_dir, "dac_enc_golden.npz"))
    sd = synth.dac_encoder_state_dict(int(g["weights_seed"]), init="test")
    e = DACVAEEncoder(precision=request.param)
    e.load_state_dict(sd)
    return g, sd, e


@pytest.mark.parametrize("case", ["a", "b"])
def test_encode_vs_reference_golden(enc, case):
    g, sd, e = enc
    audio = synth.audio_clip(int(g[f"enc_{case}_index"]), int(g[f"enc_{case}_frames"]) * 480)
    noise = torch.zeros(1, 80, int(g[f"enc_{case}_frames"]))
    z, m, logs = [t.cpu() for t in e.encode(audio.to(DEV), noise.to(DEV))]
    em = O.rel_l2(m, torch.from_numpy(g[f"enc_{case}_m"]))
    el = O.rel_l2(logs, torch.from_numpy(g[f"enc_{case}_logs"]))
    print(f"encode {case} [{e.precision}]: m rel-L2 {em:.3e}, logs rel-L2 {el:.3e}")
    assert em < TOL[e.precision] and el < TOL[e.precision]
    assert torch.equal(z, m)


def test_encode_batch_noise_and_oracle(enc):
    g, sd, e = enc
    audio = torch.cat([synth.audio_clip(5 + b, 20 * 480) for b in range(3)], 0)
    gen = torch.Generator().manual_seed(7)
    noise = torch.randn(3, 80, 20, generator=gen)
    z, m, logs = [t.cpu() for t in e.encode(audio.to(DEV), noise.to(DEV))]
    with torch.inference_mode():
        zr, mr, lr = O.dac_encode(sd, audio, noise)
    tol = TOL[e.precision]
    print(f"encode batch [{e.precision}]: m {O.rel_l2(m, mr):.3e} logs {O.rel_l2(logs, lr):.3e} z {O.rel_l2(z, zr):.3e}")
    assert O.rel_l2(m, mr) < tol and O.rel_l2(logs, lr) < tol and O.rel_l2(z, zr) < tol
    for b in range(3):  # batch == per-utterance
        z1, m1, _ = e.encode(audio[b:b + 1].to(DEV), noise[b:b + 1].to(DEV))
        assert O.rel_l2(m[b:b + 1], m1.cpu()) < 1e-5 and O.rel_l2(z[b:b + 1], z1.cpu()) < 1e-5


def test_encode_decode_round_trip_shapes_and_errors(enc):
    g, sd, e = enc
    dec = DACVAEDecoder(precision="fp32")
    dec.load_state_dict(synth.dac_decoder_state_dict(5, init="test"))
    raw = synth.audio_clip(3, 5000)  # not a multiple of the hop: preprocess pads like the reference (model.py:455-462)
    audio = e.preprocess(raw)
    assert audio.shape[-1] == 5280 and torch.equal(audio[..., :5000], raw)
    z, m, logs = e.encode(audio.to(DEV), torch.zeros(1, 80, 11).to(DEV))
    wav = dec.decode(z)
    assert wav.shape == (1, 1, 5280) and bool(torch.isfinite(wav).all())
    with torch.inference_mode():
        wav_ref = O.dac_decode(synth.dac_decoder_state_dict(5, init="test"), O.dac_encode(sd, audio)[0])
    assert O.snr_db(wav.cpu(), wav_ref) > (80.0 if e.precision == "fp32" else 30.0)
    with pytest.raises(ValueError):
        e.encode(raw.to(DEV))  # length not a multiple of the hop


def test_encode_long_batch_tensor_core_vs_oracle_crop(golden_dir):
    """16 x 10 s (the bench shape) through the tensor-core encoder; one utterance against the oracle."""
    g = np.load(os.path.join(golden_dir, "dac_enc_golden.npz"))
    sd = synth.dac_encoder_state_dict(int(g["weights_seed"]), init="test")
    e = DACVAEEncoder()
    e.load_state_dict(sd)
    audio = torch.cat([synth.audio_clip(40 + b, 240000) for b in range(16)], 0)
    z, m, logs = e.encode(audio.to(DEV), torch.zeros(16, 80, 500, device=DEV))
    assert m.shape == (16, 80, 500) and bool(torch.isfinite(m).all()) and bool(torch.isfinite(logs).all())
    with torch.inference_mode():
        _, mr, lr = O.dac_encode(sd, audio[5:6])
    em, el = O.rel_l2(m[5:6].cpu(), mr), O.rel_l2(logs[5:6].cpu(), lr)
    print(f"encode 16 x 10 s, utterance 5: m {em:.3e} logs {el:.3e}")
    assert em < 1e-2 and el < 1e-2
