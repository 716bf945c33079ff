"""Profiling driver: three launches of the DAC conv1 shape (C = 96 at 12 kHz, fp32 residual in place + Snake output)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.native as native
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_kernels_gpu as tk
DEV = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
C = int(os.environ.get("LS_C", "96"))
B, L = 16, 120000 * 96 // C
a = tk.bf16(torch.randn(B, L, C, generator=g)).to(DEV)
w1 = tk.bf16(torch.randn(1, C, C, generator=g) / math.sqrt(C)).to(DEV)
bias = (0.1 * torch.randn(C, generator=g)).to(DEV)
al = (0.5 + torch.rand(C, generator=g)).to(DEV); ia = (1.0 / (al + 1e-9))
out1 = torch.zeros(B, L, C, device=DEV, dtype=torch.bfloat16)
x = torch.randn(B, L, C, device=DEV)
for _ in range(3):
    tk.conv_gemm(a, w1, bias=bias, act=native.ACT_LRELU, addend=x, out0=x, out1=out1, out1_mode=native.OUT1_SNAKE, p1=(al, ia))
torch.cuda.synchronize()
print("ok")
