"""Development aid: in-kernel clock64 timeline of the fused transformer-block kernel (first tile of each CTA)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.native as native
import profiles.time_kernels as tk

DEV = torch.device("cuda:0")
lib = native.load()
buf = torch.zeros(148 * 256, dtype=torch.int64, device=DEV)
R = int(os.environ.get("LS_R", "16000"))
tk.tblock(R, 0)
lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8)
tk.tblock(R, 0)
lib.ls_debug_set_buffer(None, 0)
t = buf.view(148, 256).cpu()
for cta in (0, 60, 124):
    r = t[cta]
    base = int(r[0])
    names = {0: "start", 1: "P1 first slots", 2: "P1 issued", 3: "a3_ready#1", 12: "FF issued", 13: "a3_ready#2", 26: "end"}
    print(f"--- CTA {cta}  (cycles since MMA-thread start)")
    print("MMA :", " ".join(f"{names.get(i, str(i))}={int(r[i]) - base}" for i in range(0, 27) if int(r[i])))
    print("EPI :", " ".join(f"{i}={int(r[i]) - base}" for i in range(32, 56) if int(r[i])))
    print("LOAD slot-free time of loads 0..47:", [int(r[64 + i]) - base if int(r[64 + i]) else None for i in range(48)])
    print("MMA saw slots 0..15 full at:", [int(r[112 + i]) - base for i in range(16)])
    d = lambda a, b: [int(r[i]) - base if int(r[i]) else None for i in range(a, b)]
    print("MMA FF chunk 4: before AH wait, after, kb2=0 slot0/slot1, kb2=1 slot0/slot1... :", d(128, 134))
    print("MMA FF1(6): before drained wait, after, kb0..3 slot seen, committed:", d(134, 141))
    print("MMA QKV chunk 6: before drained wait, after, kb0..3 slot seen, committed:", d(144, 151))
    print("EPI FF chunk 4: before h_full, after, ld done, gelu done, ah_free ok, arrived:", d(160, 166))
    print("EPI out-proj LayerNorm (leader thread): pass 1 chunk 0 / 1 done, tmem_st drained, statistics combined, pass 2 chunk 0 / 1 done:", d(56, 62))
    print("LOAD QKV chunk 6, boxes 0..3: (slot free seen by the producer, TMA issued):", d(216, 224))
    print("MMA QKV chunk 6, kb0..3: (MMAs issued, commit issued):", d(200, 208))
    print("EPI QKV chunk 6: before h_full, after, ld+arrive, staged, bulk_wait_read, barrier, stores issued:", d(170, 177))
