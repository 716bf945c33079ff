"""Timing aid: token -> mu front half (tensor-core path and fp32 mode), 16 x 10 s (250 tokens each), and the per-kernel-kind
split of the tensor-core path (CUDA events around every launch, ls_profile_begin/end)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.native as native
import minimax_speech_b200.synth as synth
from minimax_speech_b200.front import TokenToMu
from minimax_speech_b200.speaker import LearnableSpeakerEncoder
DEV = torch.device("cuda:0")
B, T = 16, 250
toks, embs = zip(*[synth.token_inputs(b, T) for b in range(B)])
tok, emb = torch.cat(toks, 0).to(DEV), torch.cat(embs, 0).to(DEV)
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, n=10, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        scratch.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


out = {}
f = TokenToMu()
ms = timeit(lambda: f(tok, emb))
l0 = native.launch_count(); f(tok, emb); out["front_bf16_launches"] = native.launch_count() - l0
out["front_bf16_ms"] = ms
print(f"front (tensor-core) {B} x {T} tokens: {ms:.3f} ms -> {B * T / 25 / (ms / 1000):.0f} audio-s/s, {out['front_bf16_launches']} launches")
native.profile_begin()
for _ in range(5):
    f(tok, emb)
prof = native.profile_end()
out["front_bf16_kernels"] = {k: {"launches": v["launches"] / 5, "ms": v["ms"] / 5, "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] else None}
                             for k, v in prof.items() if v["launches"]}
print(json.dumps(out["front_bf16_kernels"]))
f32 = TokenToMu(precision="fp32")
ms32 = timeit(lambda: f32(tok, emb), n=2, warm=1)
out["front_fp32_ms"] = ms32
print(f"front (fp32 mode)   {B} x {T} tokens: {ms32:.2f} ms")
spk = LearnableSpeakerEncoder(precision="fp32")
mel = torch.cat([synth.reference_mel(i, 300) for i in range(B)], 0).to(DEV)
mss = timeit(lambda: spk(mel), n=3, warm=1)
out["speaker_fp32_ms"] = mss
print(f"speaker encoder (fp32 mode) {B} x 300 mel frames: {mss:.2f} ms")
spk_tc = LearnableSpeakerEncoder(precision="bf16")
mst = timeit(lambda: spk_tc(mel), n=5, warm=3)
out["speaker_bf16_ms"] = mst
print(f"speaker encoder (tensor cores) {B} x 300 mel frames: {mst:.3f} ms")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/time_front.json", "w"), indent=1)
