"""Speaker encoder and the whole ``CausalMaskedDiffWithXvec.inference`` drop-in on the GPU, fp32
mode, against golden outputs of the unmodified reference modules and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.synth as synth  # noqa: E402
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder  # noqa: E402
from minimax_speech_b200.front import CausalMaskedDiffWithXvec  # noqa: E402
from minimax_speech_b200.speaker import LearnableSpeakerEncoder  # noqa: E402
from oracle import restatement as O  # noqa: E402

DEV = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 1)


@pytest.mark.parametrize("case", ["a", "b"])
def test_speaker_embedding_vs_reference_golden(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "speaker_golden.npz"))
    sd = synth.speaker_encoder_state_dict(int(g["weights_seed"]))
    enc = LearnableSpeakerEncoder(precision="fp32")
    enc.load_state_dict(sd)
    mel = torch.cat([synth.reference_mel(i, int(g[f"spk_{case}_frames"])) for i in range(2)], 0)
    y = enc(mel.to(DEV))
    e = O.rel_l2(y.cpu(), torch.from_numpy(g[f"spk_{case}_y"]))
    print(f"speaker embedding {case} vs reference golden: rel-L2 {e:.3e}")
    assert y.shape == (2, 192) and e < 1e-4
    assert torch.allclose(y.norm(dim=1).cpu(), torch.ones(2), atol=1e-5)
    # batch == per-clip
    y0 = enc(mel[1:2].to(DEV))
    assert O.rel_l2(y0.cpu(), y[1:2].cpu()) < 1e-6


def test_several_reference_clips_vs_oracle():
    """[B, N, 80, T]: per-clip embeddings averaged and normalised again (flow.py:336-366)."""
    sd = synth.speaker_encoder_state_dict(13)
    enc = LearnableSpeakerEncoder(precision="fp32")
    enc.load_state_dict(sd)
    mels = torch.stack([torch.cat([synth.reference_mel(60 + 3 * b + i, 29) for b in range(2)], 0) for i in range(3)], dim=1)
    y = enc.encode_references(mels.to(DEV))
    with torch.inference_mode():
        ref = torch.nn.functional.normalize(torch.stack([O.speaker_encode(sd, mels[:, i]) for i in range(3)], 1).mean(1), dim=1)
    assert y.shape == (2, 192) and O.rel_l2(y.cpu(), ref) < 1e-5


@pytest.fixture(scope="module")
def pipeline():
    fsd, esd, ssd = synth.pipeline_state_dicts()
    est = CausalConditionalDecoder(precision="fp32", **synth.PIPE_EST)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    m = CausalMaskedDiffWithXvec(use_speaker_encoder=True, decoder=cfm, precision="fp32")
    full = dict(fsd)
    full.update({"decoder.estimator." + k: v for k, v in esd.items()})
    full.update({"speaker_encoder." + k: v for k, v in ssd.items()})
    m.load_state_dict(full, strict=True)
    return m


@pytest.mark.parametrize("case", ["a", "b"])
def test_inference_vs_reference_golden(golden_dir, pipeline, case):
    """The reference's own ``inference`` call (flow.py:437-511), argument for argument: prompt tokens + prompt latents, x-vector
    (a) / two reference clips through the speaker encoder, non-final chunk, streaming mask (b); 10 Euler steps with CFG."""
    g = np.load(os.path.join(golden_dir, "pipeline_golden.npz"))
    a = synth.pipeline_inputs(case)
    n = lambda t: torch.tensor([t.shape[1]], dtype=torch.int32)  # noqa: E731
    dv = lambda t: None if t is None else t.to(DEV)  # noqa: E731
    feat, none = pipeline.inference(dv(a["token"]), n(a["token"]), dv(a["prompt_token"]), n(a["prompt_token"]), dv(a["prompt_feat"]),
                                    n(a["prompt_feat"]), embedding=dv(a["embedding"]), reference_mels=dv(a["reference_mels"]),
                                    streaming=a["streaming"], finalize=a["finalize"])
    ref = torch.from_numpy(g[f"pipe_{case}_y"])
    e = O.rel_l2(feat.cpu(), ref)
    print(f"CausalMaskedDiffWithXvec.inference {case} vs reference golden: rel-L2 {e:.3e}")
    assert none is None and feat.shape == ref.shape and feat.dtype == torch.float32 and e < 1e-4


def test_inference_without_speaker_information(pipeline):
    """No x-vector and no reference mels: the zero embedding (flow.py:465-466), so spks = the affine layer's bias."""
    a = synth.pipeline_inputs("a")
    feat, _ = pipeline.inference(a["token"].to(DEV), None, a["prompt_token"].to(DEV), None, a["prompt_feat"].to(DEV), None,
                                 finalize=True)
    fsd, esd, _ = synth.pipeline_state_dicts()
    with torch.inference_mode():
        ref = O.flow_inference(fsd, esd, synth.fixed_noise(), a["token"], a["prompt_token"], a["prompt_feat"], finalize=True)
    assert O.rel_l2(feat.cpu(), ref) < 1e-4


def test_inference_tensor_core_path(golden_dir):
    """The default (bf16 tensor-core) front half + estimator against the reference's golden pipeline outputs: 1e-2 bar."""
    g = np.load(os.path.join(golden_dir, "pipeline_golden.npz"))
    fsd, esd, ssd = synth.pipeline_state_dicts()
    est = CausalConditionalDecoder(**synth.PIPE_EST)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    m = CausalMaskedDiffWithXvec(use_speaker_encoder=True, decoder=cfm)
    full = dict(fsd)
    full.update({"decoder.estimator." + k: v for k, v in esd.items()})
    full.update({"speaker_encoder." + k: v for k, v in ssd.items()})
    m.load_state_dict(full, strict=True)
    for case in ("a", "b"):
        a = synth.pipeline_inputs(case)
        dv = lambda t: None if t is None else t.to(DEV)  # noqa: E731
        feat, _ = m.inference(dv(a["token"]), None, dv(a["prompt_token"]), None, dv(a["prompt_feat"]), None, embedding=dv(a["embedding"]),
                              reference_mels=dv(a["reference_mels"]), streaming=a["streaming"], finalize=a["finalize"])
        e = O.rel_l2(feat.cpu(), torch.from_numpy(g[f"pipe_{case}_y"]))
        print(f"CausalMaskedDiffWithXvec.inference {case} (tensor-core path) vs reference golden: rel-L2 {e:.3e}")
        assert e < 1.5e-2


@pytest.mark.parametrize("case", ["a", "b"])
def test_speaker_embedding_tensor_core_path(golden_dir, case):
    """precision="bf16": conv_gemm + the flash-attention kernel (qkv rows permuted from head-major to Q | K | V), 1e-2 bar."""
    g = np.load(os.path.join(golden_dir, "speaker_golden.npz"))
    sd = synth.speaker_encoder_state_dict(int(g["weights_seed"]))
    enc = LearnableSpeakerEncoder()
    assert enc.precision == "bf16"
    enc.load_state_dict(sd)
    mel = torch.cat([synth.reference_mel(i, int(g[f"spk_{case}_frames"])) for i in range(2)], 0)
    y = enc(mel.to(DEV))
    e = O.rel_l2(y.cpu(), torch.from_numpy(g[f"spk_{case}_y"]))
    print(f"speaker embedding {case} (tensor-core path) vs reference golden: rel-L2 {e:.3e}")
    assert y.shape == (2, 192) and e < 1e-2
    assert torch.allclose(y.norm(dim=1).cpu(), torch.ones(2), atol=1e-5)
    assert torch.equal(enc(mel[1:2].to(DEV)), y[1:2])  # batch == per-clip
    mels = torch.stack([mel, mel.flip(0)], dim=1)  # [2, 2 refs, 80, T]
    ym = enc.encode_references(mels.to(DEV))
    with torch.inference_mode():
        ref = torch.nn.functional.normalize(torch.stack([O.speaker_encode(sd, mels[:, i]) for i in range(2)], 1).mean(1), dim=1)
    assert O.rel_l2(ym.cpu(), ref) < 1e-2
