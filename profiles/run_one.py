"""Profiling driver: one 1-step CFM solve (16 x 10 s, CFG) + one DAC decode, after a warm-up pass.
Used under ncu (see profiles/README.md); prints nothing that is a bench value."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.synth as synth
from minimax_speech_b200.dac import DACVAEDecoder
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder

B = int(os.environ.get("LS_B", "16"))
T = int(os.environ.get("LS_T", "500"))
dev = torch.device("cuda:0")
est = CausalConditionalDecoder()
cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
dac = DACVAEDecoder()
mu, mask, spks, cond = [t.to(dev) for t in synth.batch_inputs([T] * B)]
for it in range(2):  # pass 0 = warm-up (allocations, tensor maps); pass 1 = the one to look at
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off: only pass 1 is captured
    lat, _ = cfm(mu=mu, mask=mask, n_timesteps=1, spks=spks, cond=cond)
    wav = dac.decode(lat)
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", float(wav.abs().max()))
