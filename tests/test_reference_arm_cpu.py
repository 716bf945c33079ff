"""bench.py's reference arm: the unmodified reference modules (from /root/reference, or the copy oracle/stage_ref.py
stages under baseline/_ref) agree with the oracle restatement on the same inputs, and the staged copy is byte-identical
to its source.  Skipped where neither exists."""
import hashlib
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.reference_available(), reason="no copy of the reference on this machine")


@needs_ref
def test_reference_modules_match_the_restatement():
    import bench
    import minimax_speech_b200.synth as synth
    from oracle import restatement as O
    esd = synth.estimator_state_dict(3, init="test", n_blocks=1, num_mid_blocks=1)
    dsd = synth.dac_decoder_state_dict(5, init="test")
    # the reference estimator at the same reduced depth
    cfm = R.build_reference_flow(dict(n_blocks=1, num_mid_blocks=1))
    cfm.estimator.load_state_dict(esd, strict=True)
    dac = R.build_reference_dac()
    missing, unexpected = dac.load_state_dict(dsd, strict=False)
    assert not unexpected
    mu, mask, spks, cond = synth.batch_inputs([24])
    with torch.inference_mode():
        lat_ref, _ = cfm(mu=mu, mask=mask, n_timesteps=2, temperature=1.0, spks=spks, cond=cond)
        wav_ref = dac.decode(lat_ref)
        lat = O.cfm_forward(esd, synth.fixed_noise(), mu, mask, 2, 1.0, spks, cond)
        wav = O.dac_decode(dsd, lat)
    assert O.rel_l2(lat, lat_ref) < 1e-5
    assert O.snr_db(wav, wav_ref) > 80.0
    assert bench._ref_kind(None)[0] == "port"


@needs_ref
def test_bench_reference_runner_uses_the_reference_classes():
    import bench
    import minimax_speech_b200.synth as synth
    ref = bench.reference_modules(synth.estimator_state_dict(1986, "reference"), synth.dac_decoder_state_dict(0, "reference"))
    assert ref is not None and bench._ref_kind(ref)[0] == "reference"
    assert type(ref.cfm).__module__ == "cosyvoice.flow.flow_matching" and type(ref.cfm).__name__ == "CausalConditionalCFM"
    assert type(ref.dac).__name__ == "DACVAE"
    mod = sys.modules[type(ref.cfm).__module__]
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(R.REF_ROOT))


def test_staged_copy_is_byte_identical_to_its_source():
    man = os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")
    if not os.path.exists(man):
        pytest.skip("nothing staged")
    with open(man) as f:
        m = json.load(f)
    assert len(m["files"]) >= 10
    for rel, digest in m["files"].items():
        with open(os.path.join(ROOT, "baseline", "_ref", rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
        src = os.path.join(m["source"], rel)
        if os.path.exists(src):
            with open(src, "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel
