"""Development aid: run the bench-size flow solve and DAC decode separately with synchronisation after each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.synth as synth
from minimax_speech_b200.dac import DACVAEDecoder
from minimax_speech_b200.flow import CausalConditionalCFM, CausalConditionalDecoder
B = int(os.environ.get("LS_B", "16")); T = int(os.environ.get("LS_T", "500"))
dev = torch.device("cuda:0")
est = CausalConditionalDecoder()
cfm = CausalConditionalCFM(240, dict(t_scheduler="cosine", inference_cfg_rate=0.7), 1, 80, est)
dac = DACVAEDecoder()
mu, mask, spks, cond = [t.to(dev) for t in synth.batch_inputs([T] * B)]
for it in range(3):
    lat, _ = cfm(mu=mu, mask=mask, n_timesteps=2, spks=spks, cond=cond)
    torch.cuda.synchronize(); print("flow ok", it, float(lat.abs().max()), flush=True)
for it in range(3):
    wav = dac.decode(lat)
    torch.cuda.synchronize(); print("dac ok", it, float(wav.abs().max()), flush=True)
