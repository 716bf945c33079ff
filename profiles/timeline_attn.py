"""Development aid: in-kernel clock64 timeline of the attention kernel (first 148 CTAs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.native as native
import profiles.time_kernels as tk

DEV = torch.device("cuda:0")
lib = native.load()
buf = torch.zeros(148 * 64, dtype=torch.int64, device=DEV)
B, T = int(os.environ.get("LS_B", "32")), int(os.environ.get("LS_T", "500"))
tk.attention(B, T)
lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8)
tk.attention(B, T)
lib.ls_debug_set_buffer(None, 0)
t = buf.view(148, 64).cpu()
for cta in (0, 1, 2, 3, 100, 147):
    r = t[cta]
    base = int(r[0])
    rel = lambda i: int(r[i]) - base if int(r[i]) else None
    print(f"--- CTA {cta}: setup done {rel(15)}  ctl: loads issued {rel(1)} q/k0 landed {rel(2)}  "
          f"[bar_p seen, PV issued] x4: {[rel(i) for i in range(3, 11)]}")
    for j in range(4):
        print(f"    softmax j={j}: bar_s {rel(16+6*j)} ld {rel(17+6*j)} max {rel(18+6*j)} exp {rel(19+6*j)} "
              f"bar_o {rel(20+6*j)} stored {rel(21+6*j)}")
    print(f"    final bar_o {rel(40)} out stored {rel(41)} all done {rel(42)}")
