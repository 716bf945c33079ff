"""Experiment: does running two half-batches on two streams (tblock of one half overlapping attention of the other) beat one
batch of 16?  Two module instances (two handles: separate workspaces) with the same weights."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.synth as synth
from minimax_speech_b200.dac import DACVAEDecoder
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder
from minimax_speech_b200.pipeline import Synthesizer

dev = torch.device("cuda:0")
esd = synth.estimator_state_dict(1986, "reference")
dsd = synth.dac_decoder_state_dict(0, "reference")
n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 2


def make():
    est = CausalConditionalDecoder(); est.load_state_dict(esd)
    cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
    dac = DACVAEDecoder(); dac.load_state_dict(dsd)
    return Synthesizer(cfm, dac)


syns = [make() for _ in range(n_inst)]
B, T = 16, 500
mu, mask, spks, cond = [t.to(dev) for t in synth.batch_inputs([T] * B)]
scratch = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(n_inst)]


def step_single():
    scratch.zero_()
    return syns[0](mu, mask, spks, cond, n_timesteps=10)


def step_split(n):
    scratch.zero_()
    cur = torch.cuda.current_stream()
    outs = []
    per = B // n
    for i in range(n):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            sl = slice(i * per, (i + 1) * per)
            outs.append(syns[i](mu[sl], mask[sl], spks[sl], cond[sl], n_timesteps=10))
    for i in range(n):
        cur.wait_stream(streams[i])
    return outs


def timeit(fn, steps=10):
    for _ in range(6):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


t0 = time.time()
while time.time() - t0 < 1.5:
    step_single()
a = timeit(step_single)
print(f"single batch of 16: {a:.2f} ms/step  -> {160 / a * 1000:.0f} audio-s/s")
w0 = step_single()
for n in range(2, n_inst + 1):
    if B % n:
        continue
    b = timeit(lambda: step_split(n))
    o = torch.cat(step_split(n), 0)
    torch.cuda.synchronize()
    print(f"{n} streams x {B // n}: {b:.2f} ms/step -> {160 / b * 1000:.0f} audio-s/s; max |diff| vs single {float((o - w0).abs().max()):.2e}")
