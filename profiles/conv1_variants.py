"""Development aid: which part of the DAC conv1 epilogue (C = 96 at 12 kHz) costs the time?"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minimax_speech_b200.native as native
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_kernels_gpu as tk
DEV = torch.device("cuda:0")
def run(name, fn, gb):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 3
    print(f"{name:70s}: {us:7.0f} us  {gb / us * 1e3:6.0f} GB/s")
g = torch.Generator(device="cpu").manual_seed(0)
C = int(os.environ.get("LS_C", "96"))
B, L = 16, 120000 * 96 // C
a = tk.bf16(torch.randn(B, L, C, generator=g)).to(DEV)
w1 = tk.bf16(torch.randn(1, C, C, generator=g) / math.sqrt(C)).to(DEV)
w7 = tk.bf16(torch.randn(7, C, C, generator=g) / math.sqrt(7 * C)).to(DEV)
bias = (0.1 * torch.randn(C, generator=g)).to(DEV)
al = (0.5 + torch.rand(C, generator=g)).to(DEV); ia = (1.0 / (al + 1e-9))
out1 = torch.zeros(B, L, C, device=DEV, dtype=torch.bfloat16)
x = torch.randn(B, L, C, device=DEV)
xo = torch.zeros(B, L, C, device=DEV)
xh = torch.zeros(B, L, C, device=DEV, dtype=torch.bfloat16)
n = B * L * C / 1e9
A = native
run("conv1 bf16 copy out only", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, out1=out1, out1_mode=A.OUT1_COPY), n * 4)
run("conv1 snake bf16 out only", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, out1=out1, out1_mode=A.OUT1_SNAKE, p1=(al, ia)), n * 4)
run("conv1 f32 out only", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, out0=xo), n * 6)
run("conv1 f32 out + snake", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, out0=xo, out1=out1, out1_mode=A.OUT1_SNAKE, p1=(al, ia)), n * 8)
run("conv1 f32 addend + f32 out (separate buffers)", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, addend=x, out0=xo), n * 10)
run("conv1 f32 addend in place + snake", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, addend=x, out0=x, out1=out1, out1_mode=A.OUT1_SNAKE, p1=(al, ia)), n * 12)
run("conv1 bf16 addend + bf16 out + snake", lambda: tk.conv_gemm(a, w1, bias=bias, act=A.ACT_LRELU, addend=xh, out0=xh, out1=out1, out1_mode=A.OUT1_SNAKE, p1=(al, ia)), n * 8)
run("conv7 dil 3 snake bf16 out", lambda: tk.conv_gemm(a, w7, dil=3, pad=9, bias=bias, act=A.ACT_LRELU, out1=out1, out1_mode=A.OUT1_SNAKE, p1=(al, ia)), n * 4)
