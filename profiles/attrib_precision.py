"""Attribute the bf16 error of one estimator call: emulate operand rounding at selected places in the oracle."""
import sys, os, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
import minimax_speech_b200.synth as synth
from oracle import restatement as O
from oracle.gen_golden import est_inputs
torch.set_num_threads(os.cpu_count())
sd = synth.estimator_state_dict(7, init="test")
x, mask, mu, t, spks, cond = est_inputs([130, 77], 101)
with torch.inference_mode():
    ref = O.estimator_forward(sd, x, mask, mu, t, spks, cond)

def bf(x): return x.to(torch.bfloat16).to(torch.float32)
def hf(x): return x.to(torch.float16).to(torch.float32)
orig_linear, orig_conv, orig_matmul = F.linear, F.conv1d, torch.matmul
def run(name, rw=None, ra=None, rqk=None, rp=None, skip_time=True):
    def lin(inp, w, b=None):
        if skip_time and inp.dim() == 2:  # time MLP stays fp32 in the kernel path
            return orig_linear(inp, w, b)
        return orig_linear(ra(inp) if ra else inp, rw(w) if rw else w, b)
    def conv(inp, w, b=None, *a, **k):
        return orig_conv(ra(inp) if ra else inp, rw(w) if rw else w, b, *a, **k)
    state = {"n": 0}
    def mm(a, b):
        # transformer_block: first matmul = q k^T, second = softmax(s) v
        state["n"] += 1
        if state["n"] % 2 == 1:
            return orig_matmul(a, b)  # q, k already rounded as linear outputs? (outputs are rounded below)
        return orig_matmul(rp(a) if rp else a, b)
    F.linear, F.conv1d, torch.matmul = lin, conv, mm
    try:
        with torch.inference_mode():
            y = O.estimator_forward(sd, x, mask, mu, t, spks, cond)
    finally:
        F.linear, F.conv1d, torch.matmul = orig_linear, orig_conv, orig_matmul
    e = [O.rel_l2(y[b, :, :n], ref[b, :, :n]) for b, n in enumerate([130, 77])]
    print(f"{name:55s} rel-L2 {e[0]:.3e} {e[1]:.3e}")
run("exact")
run("weights bf16 only", rw=bf)
run("GEMM input activations bf16 only", ra=bf)
run("weights + activations bf16", rw=bf, ra=bf)
run("P (softmax probs) bf16 only", rp=bf)
run("weights + activations + P bf16", rw=bf, ra=bf, rp=bf)
run("weights fp16 only", rw=hf)
run("activations fp16 only", ra=hf)
run("weights + activations + P fp16", rw=hf, ra=hf, rp=hf)
run("weights bf16, activations fp16", rw=bf, ra=hf, rp=hf)
