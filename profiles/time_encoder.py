"""Timing aid: DAC-VAE encoder (tensor-core path), 16 x 10 s of 24 kHz audio."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.synth as synth
from minimax_speech_b200.dac import DACVAEEncoder
DEV = torch.device("cuda:0")
e = DACVAEEncoder()
B, S = 16, 240000
audio = torch.rand(B, 1, S, device=DEV) * 0.2 - 0.1
noise = torch.zeros(B, 80, S // 480, device=DEV)
for _ in range(3):
    e.encode(audio, noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    e.encode(audio, noise)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"DAC encode {B} x {S / 24000:g} s: {ms:.2f} ms  ->  {B * S / 24000 / (ms / 1000):.0f} audio-s/s")
