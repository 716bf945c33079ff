"""The drop-in boundary under the conditions SURVEY.md section 8(b) lists: work is launched on the caller's current
stream (the reference CLI runs token2wav next to an LLM thread on a side stream, cli/model.py:58,104,183), several
sessions run concurrently from different threads (one handle per thread, like one TensorRT context per thread,
common.py:171-186), errors come back as exceptions, output aliasing of the estimator seam works."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.native as native  # noqa: E402
import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.dac import DACVAEDecoder  # noqa: E402
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder  # noqa: E402
from minimax_speech_b200.pipeline import Synthesizer  # noqa: E402

DEV = torch.device("cuda:0")


def _models(seed=3):
    est = CausalConditionalDecoder(n_blocks=1, num_mid_blocks=1)
    est.load_state_dict(synth.estimator_state_dict(seed, init="test", n_blocks=1, num_mid_blocks=1))
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    dac = DACVAEDecoder()
    dac.load_state_dict(synth.dac_decoder_state_dict(5, init="test"))
    return cfm, dac


def test_side_stream_matches_default_stream():
    cfm, dac = _models()
    syn = Synthesizer(cfm, dac)
    mu, mask, spks, cond = [t.to(DEV) for t in synth.batch_inputs([90, 41])]
    ref = syn(mu, mask, spks, cond, n_timesteps=3)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):
        out = syn(mu, mask, spks, cond, n_timesteps=3)
    side.synchronize()
    assert torch.equal(out, ref)


def test_concurrent_sessions_from_threads():
    """Four threads, each with its own modules (handles) and its own stream, against the single-thread results."""
    inputs = [synth.batch_inputs([70 + 13 * i, 30 + 7 * i], first_index=10 * i) for i in range(4)]
    cfm0, dac0 = _models()
    syn0 = Synthesizer(cfm0, dac0)
    expected = [syn0(*[t.to(DEV) for t in inp], n_timesteps=3).cpu() for inp in inputs]
    results, errors = [None] * 4, []

    def work(i):
        try:
            cfm, dac = _models()
            syn = Synthesizer(cfm, dac)
            stream = torch.cuda.Stream(device=DEV)
            with torch.cuda.stream(stream):
                for _ in range(3):  # several calls per session: workspace reuse under concurrency
                    out = syn(*[t.to(DEV) for t in inputs[i]], n_timesteps=3)
            stream.synchronize()
            results[i] = out.cpu()
        except Exception as e:  # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        assert torch.equal(results[i], expected[i]), i


def test_estimator_output_may_alias_x():
    """The TensorRT branch of forward_estimator binds the output to x's buffer (flow_matching.py:136-152)."""
    cfm, _ = _models()
    from oracle.gen_golden import est_inputs
    x, mask, mu, t, spks, cond = [v.to(DEV) for v in est_inputs([50, 50], 11)]
    h = cfm.estimator.handle(DEV)
    ref = h.estimator_forward(x, mask, mu, t, spks, cond)
    xa = x.clone()
    out = h.estimator_forward(xa, mask, mu, t, spks, cond, out=xa)
    assert out.data_ptr() == xa.data_ptr() and torch.equal(out, ref)


def test_errors_are_status_codes_then_exceptions():
    lib = native.load()
    assert lib.ls_flow_solve(None, None, None, None, None, None, 0, None, 1, 1.0, 0.7, 0, None, 1, 1, None) < 0
    assert b"null" in lib.ls_last_error()
    cfm, dac = _models()
    mu, mask, spks, cond = [t.to(DEV) for t in synth.batch_inputs([40])]
    with pytest.raises((RuntimeError, ValueError)):
        cfm(mu=mu, mask=mask, n_timesteps=0, spks=spks, cond=cond)  # no steps
    with pytest.raises((RuntimeError, ValueError)):
        cfm(mu=mu[:, :40], mask=mask, n_timesteps=2, spks=spks, cond=cond)  # wrong channel count
    with pytest.raises((RuntimeError, ValueError)):
        dac.decode(torch.zeros(1, 79, 10, device=DEV))  # wrong latent dim
    bad = dict(synth.estimator_state_dict(3, init="test", n_blocks=1, num_mid_blocks=1))
    bad.pop("final_proj.bias")
    with pytest.raises(RuntimeError):
        native.FlowHandle(bad, DEV)
