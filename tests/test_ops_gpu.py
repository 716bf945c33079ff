"""The torch custom-op layer (torch.ops.ls_b200.*), CUDA-graph replay and the call-surface properties SURVEY.md section 8b
asks for: fake kernels so the ops compose (opcheck, torch.compile fullgraph), no host synchronisation on the hot call
(a bad mask is reported by the next call), one handle safely shared by several streams."""
import pytest
import torch
import torch._dynamo

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.ops  # noqa: E402,F401  (registers torch.ops.ls_b200)
import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.dac import DACVAEDecoder  # noqa: E402
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder, _t_span_values  # noqa: E402
from minimax_speech_b200.pipeline import Synthesizer  # noqa: E402

DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def models():
    est = CausalConditionalDecoder()
    est.load_state_dict(synth.estimator_state_dict(7, "test"))
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    dac = DACVAEDecoder()
    dac.load_state_dict(synth.dac_decoder_state_dict(11, "test"))
    return cfm, dac


def _inputs(lengths, first=60):
    return tuple(t.to(DEV) for t in synth.batch_inputs(lengths, first_index=first))


def test_opcheck_all_ops(models):
    cfm, dac = models
    mu, mask, spks, cond = _inputs([48, 31])
    fh, dh = cfm.estimator.handle(DEV), dac.handle(DEV)
    noise = cfm._noise_on(DEV)[0]
    t_span = list(_t_span_values(2, "cosine"))
    x = torch.randn(2, 80, 48, device=DEV)
    t = torch.rand(2, device=DEV)
    torch.library.opcheck(torch.ops.ls_b200.estimator_forward.default, (fh.key, x, mask, mu, t, spks, cond, False))
    torch.library.opcheck(torch.ops.ls_b200.flow_solve.default,
                          (fh.key, mu, mask, spks, cond, noise, t_span, 1.0, 0.7, False))
    torch.library.opcheck(torch.ops.ls_b200.mask_to_lengths.default, (mask,))
    lengths = torch.ops.ls_b200.mask_to_lengths(mask)
    assert lengths.dtype == torch.int32 and lengths.tolist() == [48, 31]
    z = torch.randn(2, 80, 48, device=DEV)
    torch.library.opcheck(torch.ops.ls_b200.dac_decode.default, (dh.key, z, lengths, dac.hop_length))
    torch.library.opcheck(torch.ops.ls_b200.dac_decode.default, (dh.key, z, None, dac.hop_length))


def test_ops_have_no_cpu_kernel(models):
    cfm, dac = models
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.ls_b200.mask_to_lengths(torch.ones(1, 1, 8))
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.ls_b200.dac_decode(dac.handle(DEV).key, torch.zeros(1, 80, 4), None, dac.hop_length)


def test_torch_compile_fullgraph_synthesizer(models):
    """One graph, three custom-op nodes: the drop-in modules trace without graph breaks and the compiled call returns
    what the eager call returns."""
    cfm, dac = models
    syn = Synthesizer(cfm, dac)
    mu, mask, spks, cond = _inputs([64, 40])
    with torch.inference_mode():
        eager = syn(mu, mask, spks, cond, n_timesteps=3)
    torch._dynamo.reset()
    seen = []

    def backend(gm, example_inputs):
        seen.append([str(n.target) for n in gm.graph.nodes if n.op == "call_function" and "ls_b200" in str(n.target)])
        return gm.forward

    fn = torch.compile(lambda a, b, c, d: syn.run(a, b, c, d, 3), backend=backend, fullgraph=True)
    with torch.no_grad():
        out = fn(mu, mask, spks, cond)
    assert len(seen) == 1 and len(seen[0]) == 3, seen
    assert torch.equal(out, eager)


def test_graph_replay_equals_eager_launches(models):
    """The CUDA-graph form of solve + decode (ls_graph_*): same bits as the eager launches, one cudaGraphLaunch."""
    import minimax_speech_b200.native as native
    cfm, dac = models
    syn = Synthesizer(cfm, dac)
    mu, mask, spks, cond = _inputs([100])
    eager = syn(mu, mask, spks, cond, n_timesteps=4).clone()
    n0 = native.launch_count()
    wav = syn.graphed(mu, mask, spks, cond, n_timesteps=4)
    assert torch.equal(wav, eager)
    g = next(iter(syn._graphs.values()))
    assert g.kernels > 500  # the whole launch sequence sits inside the graph
    # a second utterance of the same shape replays the same graph over new inputs
    mu2, mask2, spks2, cond2 = _inputs([100], first=61)
    eager2 = syn(mu2, mask2, spks2, cond2, n_timesteps=4).clone()
    n1 = native.launch_count()
    wav2 = syn.graphed(mu2, mask2, spks2, cond2, n_timesteps=4)
    assert len(syn._graphs) == 1 and torch.equal(wav2, eager2)
    assert native.launch_count() - n1 == g.kernels and n1 > n0
    # a larger eager call regrows the workspace under the graph: the next replay re-captures by itself
    big = _inputs([400, 400])
    syn(*big, n_timesteps=1)
    wav3 = syn.graphed(mu, mask, spks, cond, n_timesteps=4)
    assert torch.equal(wav3, eager)


def test_mixed_length_graph_replay(models):
    cfm, dac = models
    syn = Synthesizer(cfm, dac)
    mu, mask, spks, cond = _inputs([90, 37, 64])
    eager = syn(mu, mask, spks, cond, n_timesteps=2).clone()
    assert torch.equal(syn.graphed(mu, mask, spks, cond, n_timesteps=2), eager)


def test_non_prefix_mask_reported_by_next_call(models):
    """No host synchronisation validates the mask on the hot call; the device-side check raises at the next call."""
    cfm, _ = models
    mu, mask, spks, cond = _inputs([40])
    bad = mask.clone()
    bad[0, 0, 10] = 0.0  # a hole: not a prefix mask
    cfm(mu=mu, mask=bad, n_timesteps=1, spks=spks, cond=cond)  # returns (result undefined), flag raised on the device
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="prefix"):
        cfm(mu=mu, mask=mask, n_timesteps=1, spks=spks, cond=cond)
    y, _ = cfm(mu=mu, mask=mask, n_timesteps=1, spks=spks, cond=cond)  # the flag is cleared once reported
    assert bool(torch.isfinite(y).all())
    with pytest.raises(ValueError, match="prefix"):  # a CPU mask is still checked on the host, for free
        from minimax_speech_b200.flow import _check_prefix_mask
        _check_prefix_mask(bad.cpu())


def test_one_handle_across_streams(models):
    """A handle used on stream A and then on stream B: the second call waits for the first (shared workspace)."""
    cfm, _ = models
    mu, mask, spks, cond = _inputs([120, 120])
    ref, _ = cfm(mu=mu, mask=mask, n_timesteps=3, spks=spks, cond=cond)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for s in (sa, sb, sa, sb):
        with torch.cuda.stream(s):
            outs.append(cfm(mu=mu, mask=mask, n_timesteps=3, spks=spks, cond=cond)[0])
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, ref)


def test_growth_is_stream_ordered(models):
    """Growing the workspace (a larger shape arrives) needs no device-wide synchronisation: results before and after are
    the same bits."""
    cfm, dac = models
    est2 = CausalConditionalDecoder()
    est2.load_state_dict(synth.estimator_state_dict(7, "test"))
    cfm2 = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est2)
    small = _inputs([50])
    big = _inputs([300, 300, 300])
    a1, _ = cfm2(mu=small[0], mask=small[1], n_timesteps=2, spks=small[2], cond=small[3])
    b1, _ = cfm2(mu=big[0], mask=big[1], n_timesteps=2, spks=big[2], cond=big[3])  # regrows, no sync
    a2, _ = cfm2(mu=small[0], mask=small[1], n_timesteps=2, spks=small[2], cond=small[3])
    ref_b, _ = cfm(mu=big[0], mask=big[1], n_timesteps=2, spks=big[2], cond=big[3])
    assert torch.equal(a1, a2) and torch.equal(b1, ref_b)
