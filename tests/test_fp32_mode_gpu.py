"""fp32 mode (precision="fp32": fp32 end to end on the CUDA cores) against the golden outputs of the unmodified
reference and the CPU oracle.  Tolerance is north_star's: latent rel-L2 <= 1e-4; the waveform bound that corresponds
to it (SNR = -20 log10(rel-L2)) is 80 dB."""
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
from oracle import restatement as O  # noqa: E402
from oracle.gen_golden import est_inputs  # noqa: E402

DEV = torch.device("cuda:0")
FP32_TOL = 1e-4
FP32_SNR_DB = 80.0
torch.set_num_threads(os.cpu_count() or 1)


@pytest.fixture(scope="module")
def flow(golden_dir):
    g = np.load(os.path.join(golden_dir, "flow_golden.npz"))
    sd = synth.estimator_state_dict(int(g["weights_seed"]), init="test")
    est = CausalConditionalDecoder(precision="fp32")
    est.load_state_dict(sd)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    return g, sd, cfm


@pytest.fixture(scope="module")
def dac(golden_dir):
    g = np.load(os.path.join(golden_dir, "dac_golden.npz"))
    sd = synth.dac_decoder_state_dict(int(g["weights_seed"]), init="test")
    dec = DACVAEDecoder(precision="fp32")
    dec.load_state_dict(sd)
    return g, sd, dec


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_fp32_estimator_vs_reference_golden(flow, case):
    g, sd, cfm = flow
    lengths = [int(v) for v in g[f"est_{case}_lengths"]]
    x, mask, mu, t, spks, cond = est_inputs(lengths, int(g[f"est_{case}_seed"]))
    y = cfm.forward_estimator(x.to(DEV), mask.to(DEV), mu.to(DEV), t.to(DEV), spks.to(DEV), cond.to(DEV),
                              streaming=bool(g[f"est_{case}_streaming"])).cpu()
    ref = torch.from_numpy(g[f"est_{case}_y"])
    for b, n in enumerate(lengths):
        e = O.rel_l2(y[b, :, :n], ref[b, :, :n])
        print(f"fp32 estimator {case}[{b}] rel-L2 {e:.3e}")
        assert e < FP32_TOL
        assert float(y[b, :, n:].abs().max() if n < y.shape[2] else 0.0) == 0.0


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_fp32_cfm_solve_vs_reference_golden(flow, case):
    g, sd, cfm = flow
    lengths = [int(v) for v in g[f"cfm_{case}_lengths"]]
    mu, mask, spks, cond = synth.batch_inputs(lengths, first_index=50)
    y, _ = cfm(mu=mu.to(DEV), mask=mask.to(DEV), n_timesteps=int(g[f"cfm_{case}_steps"]), temperature=1.0,
               spks=spks.to(DEV), cond=cond.to(DEV), streaming=bool(g[f"cfm_{case}_streaming"]))
    y = y.cpu()
    ref = torch.from_numpy(g[f"cfm_{case}_y"])
    for b, n in enumerate(lengths):
        e = O.rel_l2(y[b, :, :n], ref[b, :, :n])
        print(f"fp32 cfm {case}[{b}] rel-L2 {e:.3e}")
        assert e < FP32_TOL
        assert float(y[b, :, n:].abs().max() if n < y.shape[2] else 0.0) == 0.0


def test_fp32_cfm_10_steps_vs_oracle(flow):
    g, sd, cfm = flow
    mu, mask, spks, cond = synth.batch_inputs([150], first_index=7)
    y, _ = cfm(mu=mu.to(DEV), mask=mask.to(DEV), n_timesteps=10, spks=spks.to(DEV), cond=cond.to(DEV))
    with torch.inference_mode():
        ref = O.cfm_forward(sd, synth.fixed_noise(), mu, mask, 10, 1.0, spks, cond)
    e = O.rel_l2(y.cpu(), ref)
    print(f"fp32 cfm 10-step rel-L2 {e:.3e}")
    assert e < FP32_TOL


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_fp32_dac_decode_vs_reference_golden(dac, case):
    g, sd, dec = dac
    z = synth.dac_latents(int(g[f"dac_{case}_index"]), int(g[f"dac_{case}_frames"]))
    y = dec.decode(z.to(DEV)).cpu()
    ref = torch.from_numpy(g[f"dac_{case}_y"])
    assert y.shape == ref.shape
    s = O.snr_db(y, ref)
    print(f"fp32 dac {case} SNR {s:.1f} dB")
    assert s > FP32_SNR_DB


def test_fp32_dac_varlen_and_end_to_end(flow, dac):
    g, sd, cfm = flow
    _, dsd, dec = dac
    lengths = [60, 33]
    mu, mask, spks, cond = synth.batch_inputs(lengths, first_index=3)
    lat, _ = cfm(mu=mu.to(DEV), mask=mask.to(DEV), n_timesteps=3, spks=spks.to(DEV), cond=cond.to(DEV))
    wav = dec.decode(lat, torch.tensor(lengths)).cpu()
    with torch.inference_mode():
        lat_ref = O.cfm_forward(sd, synth.fixed_noise(), mu, mask, 3, 1.0, spks, cond) * mask
        wav_ref = O.dac_decode_varlen(dsd, lat_ref, lengths)
    e = O.rel_l2(lat.cpu(), lat_ref)
    s = O.snr_db(wav, wav_ref)
    print(f"fp32 end to end: latent rel-L2 {e:.3e}, waveform SNR {s:.1f} dB")
    assert e < FP32_TOL and s > FP32_SNR_DB
    assert float(wav[1, :, 33 * dec.hop_length:].abs().max()) == 0.0
