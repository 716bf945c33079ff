"""Development aid: in-kernel clock64 timeline of the fused transformer-block kernel (first tile of each CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.native as native
import profiles.time_kernels as tk

DEV = torch.device("cuda:0")
lib = native.load()
buf = torch.zeros(148 * 128, dtype=torch.int64, device=DEV)
R = int(os.environ.get("LS_R", "16000"))
tk.tblock(R, 0)
lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8)
tk.tblock(R, 0)
lib.ls_debug_set_buffer(None, 0)
t = buf.view(148, 128).cpu()
for cta in (0, 60, 124):
    r = t[cta]
    base = int(r[0])
    names = {0: "start", 1: "P1 first slots", 2: "P1 issued", 3: "a3_ready#1", 12: "FF issued", 13: "a3_ready#2", 26: "end"}
    print(f"--- CTA {cta}  (cycles since MMA-thread start)")
    print("MMA :", " ".join(f"{names.get(i, str(i))}={int(r[i]) - base}" for i in range(0, 27) if int(r[i])))
    print("EPI :", " ".join(f"{i}={int(r[i]) - base}" for i in range(32, 56) if int(r[i])))
    print("LOAD slot-free time of loads 0..47:", [int(r[64 + i]) - base if int(r[64 + i]) else None for i in range(48)])
    print("MMA saw slots 0..15 full at:", [int(r[112 + i]) - base for i in range(16)])
