"""Development aid: per-SM schedule of the attention kernel's CTAs (smid, start, first MMA, exit clocks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import minimax_speech_b200.native as native
import profiles.time_kernels as tk

DEV = torch.device("cuda:0")
lib = native.load()
B, T = int(os.environ.get("LS_B", "32")), int(os.environ.get("LS_T", "500"))
n_cta = ((T + 127) // 128) * 8 * B
buf = torch.zeros(148 * 64 + 4 * n_cta, dtype=torch.int64, device=DEV)
tk.attention(B, T)
lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8)
tk.attention(B, T)
lib.ls_debug_set_buffer(None, 0)
t = buf[148 * 64:].view(n_cta, 4).cpu()
by_sm = {}
for i in range(n_cta):
    sm, a, m, e = (int(v) for v in t[i])
    by_sm.setdefault(sm, []).append((a, m, e, i))
spans = []
for sm in sorted(by_sm)[:4] + sorted(by_sm)[-2:]:
    rows = sorted(by_sm[sm])
    base = rows[0][0]
    print(f"SM {sm}: {len(rows)} CTAs:", " ".join(f"[{a - base}+{m - a}>{e - base}]" for a, m, e, _ in rows))
for sm, rows in by_sm.items():
    rows = sorted(rows)
    spans.append((max(r[2] for r in rows) - rows[0][0], len(rows)))
print("per-SM busy span (clk): min %d  median %d  max %d;  CTAs per SM min %d max %d" % (
    min(s for s, _ in spans), sorted(s for s, _ in spans)[len(spans) // 2], max(s for s, _ in spans),
    min(n for _, n in spans), max(n for _, n in spans)))
life = sorted(int(t[i, 3] - t[i, 1]) for i in range(n_cta))
print("CTA lifetime (clk): min %d median %d p90 %d max %d" % (life[0], life[len(life) // 2], life[int(len(life) * 0.9)], life[-1]))
