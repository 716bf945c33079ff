"""DAC decoder alone: device time of one decode (16 x 10 s and 8 x 30 s latents) and the waveform of a fixed input, saved so
that two library builds (LS_LIB / LS_BUILD_DEFINES) can be compared on the same box.  Usage: time_dac.py <tag>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.synth as synth
from minimax_speech_b200.dac import DACVAEDecoder

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
dev = torch.device("cuda:0")
for init in ("reference", "trained"):
    dac = DACVAEDecoder()
    dac.load_state_dict(synth.dac_decoder_state_dict(0, init))
    z = torch.cat([synth.dac_latents(b, 500) for b in range(4)], 0).to(dev)
    wav = dac.decode(z)
    torch.save(wav.cpu(), f"/tmp/dacwav_{tag}_{init}.pt")
    d32 = DACVAEDecoder(precision="fp32")
    d32.load_state_dict(synth.dac_decoder_state_dict(0, init))
    ref = d32.decode(z)
    err = (wav - ref).double()
    snr = 10 * torch.log10(ref.double().pow(2).sum() / err.pow(2).sum())
    print(f"{tag} init={init}: SNR vs the fp32 device path {float(snr):.2f} dB, max |wav| {float(wav.abs().max()):.3f}")
    del d32
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B, L in ((16, 500), (8, 1500)):
    z = torch.cat([synth.dac_latents(b, L) for b in range(B)], 0).to(dev)
    for _ in range(3):
        dac.decode(z)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record()
        for _ in range(4):
            dac.decode(z)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 4)
    print(f"{tag} decode {B} x {L} frames: {min(ts):.3f} ms (min of 5 x 4), median {sorted(ts)[2]:.3f}")
