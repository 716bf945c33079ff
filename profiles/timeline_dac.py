"""Development aid: steady-state per-tile timeline of conv_gemm for DAC-shaped launches (stage 4: C = 96 at 12 kHz)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.native as native
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_kernels_gpu as tk
DEV = torch.device("cuda:0")
lib = native.load()
buf = torch.zeros(148 * 64, dtype=torch.int64, device=DEV)
def run(name, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000
    buf.zero_(); lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8); fn(); torch.cuda.synchronize(); lib.ls_debug_set_buffer(None, 0)
    r = buf.view(148, 64).cpu()[77]; base = int(r[0])
    rel = lambda i: int(r[i]) - base if int(r[i]) else None
    print(f"=== {name}: {us:.0f} us\n  producer tile starts {[rel(56+i) for i in range(8)]}\n  MMA tile commits     {[rel(24+i) for i in range(8)]}\n  epilogue tile done   {[rel(48+i) for i in range(8)]}\n  MMA saw B box of k-iter 0..15 at {[rel(8+i) for i in range(16)]} setup {rel(1)} pdl {rel(2)}")
    if os.environ.get("LS_DETAIL"):
        print("  detail triplets (before wait, after wait, after issue) x 7:", [[rel(8 + 3 * i + j) for j in range(3)] for i in range(7)])
g = torch.Generator(device="cpu").manual_seed(0)
C = int(os.environ.get('LS_C', '96'))
B, L = 16, 120000 * 96 // C   # 10 s at 12 kHz (C = 96) / 24 kHz (C = 48)
a = tk.bf16(torch.randn(B, L, C, generator=g)).to(DEV)
w7 = tk.bf16(torch.randn(7, C, C, generator=g) / math.sqrt(7 * C)).to(DEV)
w1 = tk.bf16(torch.randn(1, C, C, generator=g) / math.sqrt(C)).to(DEV)
bias = (0.1 * torch.randn(C, generator=g)).to(DEV)
al = (0.5 + torch.rand(C, generator=g)).to(DEV); ia = (1.0 / (al + 1e-9))
out1 = torch.zeros(B, L, C, device=DEV, dtype=torch.bfloat16)
x = torch.randn(B, L, C, device=DEV)
for dil in (1, 9):
    run(f"conv7 dil {dil} : lrelu + snake bf16 out", lambda: tk.conv_gemm(a, w7, dil=dil, pad=3 * dil, bias=bias, act=native.ACT_LRELU, out1=out1, out1_mode=native.OUT1_SNAKE, p1=(al, ia)))
run("conv1 : lrelu + x residual f32 in/out + snake bf16 out", lambda: tk.conv_gemm(a, w1, bias=bias, act=native.ACT_LRELU, addend=x, out0=x, out1=out1, out1_mode=native.OUT1_SNAKE, p1=(al, ia)))
