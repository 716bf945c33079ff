"""Development aid: in-kernel clock64 timeline of conv_gemm for estimator-shaped launches (first tile of each CTA)."""
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
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    buf.zero_()
    lib.ls_debug_set_buffer(native.ptr(buf), buf.numel() * 8)
    fn(); torch.cuda.synchronize()
    lib.ls_debug_set_buffer(None, 0)
    t = buf.view(148, 64).cpu()
    print(f"=== {name}: {us:.1f} us per launch")
    for cta in (0, 77):
        r = t[cta]; base = int(r[0])
        rel = lambda i: int(r[i]) - base if int(r[i]) else None
        print(f"  CTA {cta}: setup {rel(1)} pdl_wait {rel(2)} | MMA saw k-iter data at {[rel(8+i) for i in range(24) if int(r[8+i])]} | tile committed {rel(32)}"
              f" | EPI: acc ready {rel(40)} regs loaded {rel(41)} bias/act done {rel(45)} LN done {rel(46)} chunk starts {[rel(47+i) for i in range(4)]} main stores done {rel(42)} epilogue done {rel(43)} exit {rel(44)}")

    if os.environ.get("LS_DETAIL") == "3":
        r = t[0]; base = int(r[0])
        print("  pass B, chunks 0..3 of one epilogue thread: [start, acc loaded, LN+Mish(+temb) done, residual added, out0 store issued]:",
              [[int(r[8 + 6 * k + j]) - base if int(r[8 + 6 * k + j]) else None for j in range(5)] for k in range(4)])

g = torch.Generator(device="cpu").manual_seed(0)
B, T = 32, 500
def mk(*s, scale=1.0): return (torch.randn(*s, generator=g) * scale)
# conv2 of a resnet: K=256, taps 3, LN+Mish + addend + out0 f32 + LN out1
a = tk.bf16(mk(B, T, 256)).to(DEV); w = tk.bf16(mk(3, 256, 256, scale=1 / math.sqrt(768))).to(DEV)
bias = mk(256, scale=0.1).to(DEV); lg = (1 + 0.1 * mk(256)).to(DEV); lb = mk(256, scale=0.1).to(DEV)
addend = mk(B, T, 256).to(DEV); out0 = torch.zeros(B, T, 256, device=DEV); out1 = torch.zeros(B, T, 256, device=DEV, dtype=torch.bfloat16)
run("conv2 (k=3, 256->256, LN+Mish, +res, f32 out + LN bf16 out)", lambda: tk.conv_gemm(a, w, pad=2, bias=bias, act=native.ACT_LN_MISH, ln=(lg, lb), addend=addend, out0=out0, out1=out1, out1_mode=native.OUT1_LN, p1=(lg, lb)))
# res conv 1x1 256->256 f32 out
w1 = tk.bf16(mk(1, 256, 256, scale=1 / 16)).to(DEV)
run("res_conv (1x1, 256->256, f32 out)", lambda: tk.conv_gemm(a, w1, bias=bias, out0=out0))
# QKV: 256 -> 1536 bf16 out
wq = tk.bf16(mk(1, 1536, 256, scale=1 / 16)).to(DEV); q = torch.zeros(B, T, 1536, device=DEV, dtype=torch.bfloat16)
run("qkv (256->1536, bf16 out)", lambda: tk.conv_gemm(a, wq, block_n=256, out1=q, out1_mode=native.OUT1_COPY))
run("res_conv, no outputs at all (bias only)", lambda: tk.conv_gemm(a, w1, bias=bias))
run("res_conv, no bias, f32 out", lambda: tk.conv_gemm(a, w1, out0=out0))
out0h = torch.zeros(B, T, 256, device=DEV, dtype=torch.bfloat16)
run("res_conv, no bias, bf16 out0", lambda: tk.conv_gemm(a, w1, out0=out0h))
