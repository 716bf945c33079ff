"""Kernel-level parity on the B200: each sm_100a kernel, called through the C-ABI test hooks, against the
same op written in plain fp32 torch on the same (bf16-rounded) operands."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import minimax_speech_b200.native as native  # noqa: E402

DEV = torch.device("cuda:0")


def bf16(t):
    return t.to(torch.bfloat16)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def conv_gemm(a0, w, *, a1=None, M=None, block_n=None, dil=1, pad=0, lengths=None, m_len_mul=1, m_len_add=0,
              skip_halo=0, chan_mod=None, bias=None, act=0, ln=None, temb=None, temb_bstride=0, addend=None,
              out0=None, out1=None, out1_mode=0, p1=None, n_store=None, out_ld=None, out_shift=0, out_bstride=None,
              out_alloc=None, out_valid_mul=None):
    """a0 [B,T,C0] bf16, w [taps,N,K] bf16."""
    lib = native.load()
    B, T_in, C0 = a0.shape
    taps, N, K = w.shape
    d = native.ConvGemmDesc()
    d.a0, d.a0_C, d.T_in = native.ptr(a0), C0, T_in
    d.a1, d.a1_C = (native.ptr(a1), a1.shape[2]) if a1 is not None else (C.c_void_p(0), 0)
    d.w, d.K = native.ptr(w), K
    d.B, d.M, d.N = B, M or T_in, N
    d.block_n = block_n or min(N, 256)
    d.taps, d.dil, d.pad = taps, dil, pad
    d.lengths = native.ptr(lengths)
    d.m_len_mul, d.m_len_add, d.skip_halo = m_len_mul, m_len_add, skip_halo
    d.chan_mod = chan_mod or N
    d.bias = native.ptr(bias)
    d.act = act
    d.ln_g, d.ln_b = (native.ptr(ln[0]), native.ptr(ln[1])) if ln else (C.c_void_p(0), C.c_void_p(0))
    d.temb, d.temb_bstride = native.ptr(temb), temb_bstride
    d.addend = native.ptr(addend)
    d.addend_dtype = 0 if addend is None else (1 if addend.dtype == torch.float32 else 2)
    d.out0 = native.ptr(out0)
    d.out0_dtype = 0 if out0 is None else (1 if out0.dtype == torch.float32 else 2)
    d.out1, d.out1_mode = native.ptr(out1), out1_mode
    d.p1_a, d.p1_b = (native.ptr(p1[0]), native.ptr(p1[1])) if p1 else (C.c_void_p(0), C.c_void_p(0))
    d.n_store = n_store or N
    d.out_ld = out_ld or N
    d.out_shift = out_shift
    rows = d.M
    d.out_bstride = out_bstride if out_bstride is not None else rows * d.out_ld
    d.out_alloc = out_alloc if out_alloc is not None else rows * d.out_ld
    d.out_valid_mul = out_valid_mul if out_valid_mul is not None else d.out_ld
    native.check(lib.ls_test_conv_gemm(C.byref(d), native.current_stream_ptr(DEV)), "ls_test_conv_gemm")
    torch.cuda.synchronize()


def ref_conv(a, w, dil, pad):
    """a [B,T,C] , w [taps,N,K] -> [B,T,N] with rows outside [0,T) = 0, tap k reads row t + k*dil - pad."""
    B, T, Cin = a.shape
    taps, N, K = w.shape
    x = a.float().transpose(1, 2)
    right = max(0, (taps - 1) * dil - pad)
    x = F.pad(x, (pad, right))
    wt = w.float().permute(1, 2, 0).contiguous()  # [N,K,taps]
    y = F.conv1d(x, wt, dilation=dil)
    return y[:, :, :T].transpose(1, 2).contiguous()


def mish(x):
    return x * torch.tanh(F.softplus(x))


def test_linear_bias_f32():
    g = torch.Generator(device="cpu").manual_seed(1)
    B, T, K, N = 2, 300, 256, 256
    a = bf16(torch.randn(B, T, K, generator=g)).to(DEV)
    w = bf16(torch.randn(1, N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    out = torch.full((B, T, N), float("nan"), device=DEV)
    conv_gemm(a, w, bias=bias, out0=out)
    ref = ref_conv(a, w, 1, 0) + bias
    assert rel(out, ref) < 1e-5, rel(out, ref)


@pytest.mark.parametrize("N,block_n,K", [(1536, 256, 256), (1024, 256, 256), (256, 256, 1024), (80, 80, 256),
                                         (768, 256, 64)])
def test_linear_shapes_bf16_out(N, block_n, K):
    g = torch.Generator(device="cpu").manual_seed(2)
    B, T = 3, 200
    a = bf16(torch.randn(B, T, K, generator=g)).to(DEV)
    w = bf16(torch.randn(1, N, K, generator=g) / math.sqrt(K)).to(DEV)
    out1 = torch.zeros(B, T, N, device=DEV, dtype=torch.bfloat16)
    conv_gemm(a, w, block_n=block_n, out1=out1, out1_mode=native.OUT1_COPY)
    ref = ref_conv(a, w, 1, 0)
    assert rel(out1.float(), ref) < 4e-3, rel(out1.float(), ref)


def test_causal_conv3_ln_mish_temb_addend_ln2():
    g = torch.Generator(device="cpu").manual_seed(3)
    B, T, K, N = 2, 333, 320, 256
    a = bf16(torch.randn(B, T, K, generator=g)).to(DEV)
    w = bf16(torch.randn(3, N, K, generator=g) / math.sqrt(3 * K)).to(DEV)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    lg, lb = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    g2, b2 = (1 + 0.1 * torch.randn(N, generator=g)).to(DEV), (0.1 * torch.randn(N, generator=g)).to(DEV)
    temb = torch.randn(B, N, generator=g).to(DEV)
    addend = torch.randn(B, T, N, generator=g).to(DEV)
    out0 = torch.zeros(B, T, N, device=DEV)
    out1 = torch.zeros(B, T, N, device=DEV, dtype=torch.bfloat16)
    conv_gemm(a, w, pad=2, bias=bias, act=native.ACT_LN_MISH, ln=(lg, lb), temb=temb, temb_bstride=N,
              addend=addend, out0=out0, out1=out1, out1_mode=native.OUT1_LN, p1=(g2, b2))
    y = ref_conv(a, w, 1, 2) + bias
    y = mish(F.layer_norm(y, (N,), lg, lb, 1e-5)) + temb[:, None, :] + addend
    assert rel(out0, y) < 2e-5, rel(out0, y)
    y1 = F.layer_norm(y, (N,), g2, b2, 1e-5)
    assert rel(out1.float(), y1) < 4e-3, rel(out1.float(), y1)


def test_gelu_epilogue():
    g = torch.Generator(device="cpu").manual_seed(4)
    B, T, K, N = 1, 257, 256, 1024
    a = bf16(torch.randn(B, T, K, generator=g)).to(DEV)
    w = bf16(torch.randn(1, N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = (0.1 * torch.randn(N, generator=g)).to(DEV)
    out0 = torch.zeros(B, T, N, device=DEV)
    conv_gemm(a, w, bias=bias, act=native.ACT_GELU, out0=out0)
    ref = F.gelu(ref_conv(a, w, 1, 0) + bias)
    assert rel(out0, ref) < 1e-5, rel(out0, ref)


def test_k_split_concat_and_residual_inplace():
    g = torch.Generator(device="cpu").manual_seed(5)
    B, T, N = 2, 140, 256
    a0 = bf16(torch.randn(B, T, 256, generator=g)).to(DEV)
    a1 = bf16(torch.randn(B, T, 256, generator=g)).to(DEV)
    w = bf16(torch.randn(3, N, 512, generator=g) / math.sqrt(3 * 512)).to(DEV)
    u = torch.randn(B, T, N, generator=g).to(DEV)
    u0 = u.clone()
    conv_gemm(a0, w, a1=a1, pad=2, addend=u, out0=u)
    ref = ref_conv(torch.cat([a0, a1], -1), w, 1, 2) + u0
    assert rel(u, ref) < 1e-5, rel(u, ref)


def test_lengths_mask_and_tile_skip():
    g = torch.Generator(device="cpu").manual_seed(6)
    B, T, K, N = 3, 400, 256, 256
    lengths = torch.tensor([400, 130, 7], dtype=torch.int32, device=DEV)
    a = bf16(torch.randn(B, T, K, generator=g)).to(DEV)
    w = bf16(torch.randn(1, N, K, generator=g) / math.sqrt(K)).to(DEV)
    out0 = torch.full((B, T, N), 7.0, device=DEV)
    conv_gemm(a, w, lengths=lengths, out0=out0)
    ref = ref_conv(a, w, 1, 0)
    for b, n in enumerate(lengths.tolist()):
        assert rel(out0[b, :n], ref[b, :n]) < 1e-5
        last_tile_end = min(T, (n + 127) // 128 * 128)
        assert float(out0[b, n:last_tile_end].abs().max() if last_tile_end > n else 0.0) == 0.0
        if last_tile_end < T:  # skipped tiles stay untouched
            assert float((out0[b, last_tile_end:] - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("C_,dil", [(96, 9), (48, 3), (192, 1), (768, 9)])
def test_dilated_conv7_lrelu_snake(C_, dil):
    g = torch.Generator(device="cpu").manual_seed(7)
    B, T = 2, 500
    a = bf16(torch.randn(B, T, C_, generator=g)).to(DEV)
    w = bf16(torch.randn(7, C_, C_, generator=g) / math.sqrt(7 * C_)).to(DEV)
    bias = (0.1 * torch.randn(C_, generator=g)).to(DEV)
    alpha = (0.3 * torch.randn(C_, generator=g)).to(DEV)
    ialpha = 1.0 / (alpha + 1e-9)
    out1 = torch.zeros(B, T, C_, device=DEV, dtype=torch.bfloat16)
    bn = C_ if C_ <= 256 else 256
    conv_gemm(a, w, block_n=bn, dil=dil, pad=3 * dil, bias=bias, act=native.ACT_LRELU, out1=out1,
              out1_mode=native.OUT1_SNAKE, p1=(alpha, ialpha), skip_halo=64)
    y = F.leaky_relu(ref_conv(a, w, dil, 3 * dil) + bias, 0.1)
    y = y + ialpha * torch.sin(alpha * y) ** 2
    assert rel(out1.float(), y) < 4e-3, rel(out1.float(), y)


@pytest.mark.parametrize("cin,cout,s", [(96, 48, 2), (192, 96, 3), (384, 192, 4), (1536, 768, 5)])
def test_conv_transpose_polyphase(cin, cout, s):
    g = torch.Generator(device="cpu").manual_seed(8)
    B, L = 2, 150
    pad = math.ceil(s / 2)
    x = bf16(torch.randn(B, L, cin, generator=g)).to(DEV)
    wt = bf16(torch.randn(cin, cout, 2 * s, generator=g) / math.sqrt(cin)).to(DEV)  # ConvTranspose1d layout
    bias = (0.1 * torch.randn(cout, generator=g)).to(DEV)
    # polyphase packing (see dac_engine.cu pack_wn_convT)
    w = torch.zeros(2, s * cout, cin, device=DEV, dtype=torch.bfloat16)
    for phi in range(s):
        w[1, phi * cout:(phi + 1) * cout] = wt[:, :, phi].t()
        w[0, phi * cout:(phi + 1) * cout] = wt[:, :, phi + s].t()
    rows_out = L * s
    out0 = torch.full((B, rows_out, cout), float("nan"), device=DEV)
    N = s * cout
    bn = max(b for b in range(16, 257, 16) if N % b == 0)
    conv_gemm(x, w, M=L + 1, block_n=bn, pad=1, chan_mod=cout, bias=bias, out0=out0, out_ld=N,
              out_shift=-pad * cout, out_bstride=rows_out * cout, out_alloc=rows_out * cout,
              out_valid_mul=rows_out * cout)
    ref = F.conv_transpose1d(x.float().transpose(1, 2), wt.float(), bias, stride=s, padding=pad,
                             output_padding=s % 2).transpose(1, 2)
    assert ref.shape == out0.shape
    assert rel(out0, ref) < 1e-5, rel(out0, ref)


def test_final_conv_tanh_mono():
    g = torch.Generator(device="cpu").manual_seed(9)
    B, T, C_ = 2, 1000, 48
    a = bf16(torch.randn(B, T, C_, generator=g)).to(DEV)
    w1 = bf16(torch.randn(7, 1, C_, generator=g) / math.sqrt(7 * C_)).to(DEV)
    w = torch.zeros(7, 16, C_, device=DEV, dtype=torch.bfloat16)
    w[:, :1] = w1
    bias = torch.zeros(16, device=DEV)
    bias[0] = 0.05
    wav = torch.full((B, T), float("nan"), device=DEV)
    conv_gemm(a, w, block_n=16, dil=1, pad=3, chan_mod=16, bias=bias, act=native.ACT_LRELU_TANH, out0=wav,
              n_store=1, out_ld=1, out_bstride=T, out_alloc=T, out_valid_mul=T)
    ref = torch.tanh(F.leaky_relu(ref_conv(a, w1, 1, 3)[..., 0] + 0.05, 0.1))
    assert rel(wav, ref) < 1e-5, rel(wav, ref)


def ref_attention(qkv, lengths, H, chunk):
    B, T, _ = qkv.shape
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) * 0.125
    pos = torch.arange(T, device=qkv.device)
    key_ok = pos[None, :] < lengths[:, None]                      # [B,T]
    m = key_ok[:, None, None, :].expand(B, 1, T, T)
    if chunk:
        m = m & (pos[None, :] < ((pos // chunk + 1) * chunk)[:, None])[None, None]
    s = s.masked_fill(~m, float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, T, H * 64)


@pytest.mark.parametrize("T,lengths,chunk", [(128, [128], 0), (500, [500, 377], 0), (300, [300, 40, 129], 50),
                                             (1500, [1500], 0)])
def test_attention(T, lengths, chunk):
    g = torch.Generator(device="cpu").manual_seed(10)
    B, H = len(lengths), 8
    qkv = bf16(torch.randn(B, T, 3 * H * 64, generator=g)).to(DEV)
    ln = torch.tensor(lengths, dtype=torch.int32, device=DEV)
    out = torch.zeros(B, T, H * 64, device=DEV, dtype=torch.bfloat16)
    native.check(native.load().ls_test_attention(native.ptr(qkv), native.ptr(out), native.ptr(ln), B, T, H, chunk,
                                                 native.current_stream_ptr(DEV)), "ls_test_attention")
    torch.cuda.synchronize()
    ref = ref_attention(qkv, ln, H, chunk)
    for b, n in enumerate(lengths):
        e = rel(out[b, :n].float(), ref[b, :n])
        assert e < 1e-2, (b, e)


# ------------------------------------------------------------------------------------------ fused transformer block
def _tblock_operands(R, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)
    att = bf16(rn(R, 512))
    u = rn(R, 256, scale=2.0) + 0.5
    wo = bf16(rn(256, 512, scale=512 ** -0.5))
    w1 = bf16(rn(1024, 256, scale=256 ** -0.5))
    w2 = bf16(rn(256, 1024, scale=1024 ** -0.5))
    wqkv = bf16(rn(1536, 256, scale=256 ** -0.5))
    vec = torch.cat([rn(256, scale=0.1), 1 + rn(256, scale=0.1), rn(256, scale=0.1), rn(1024, scale=0.2),
                     rn(256, scale=0.1), 1 + rn(256, scale=0.1), rn(256, scale=0.1)]).contiguous()
    return att, u, wo, w1, w2, wqkv, vec


def _tblock_ref(att, u, wo, w1, w2, wqkv, vec):
    bo, g3, be3, b1, b2, g1n, be1n = torch.split(vec, [256, 256, 256, 1024, 256, 256, 256])
    u1 = att.float() @ wo.float().T + bo + u
    n3 = bf16(F.layer_norm(u1, (256,), g3, be3, 1e-5)).float()
    h = bf16(F.gelu(n3 @ w1.float().T + b1)).float()
    u2 = u1 + h @ w2.float().T + b2
    n1 = bf16(F.layer_norm(u2, (256,), g1n, be1n, 1e-5)).float()
    return u2, n1 @ wqkv.float().T


def _tblock(att, u, wo, w1, w2, wqkv, vec, tail_mode, lengths=None, T=None):
    R = att.shape[0]
    qkv = torch.full((R, 1536), float("nan"), device=DEV, dtype=torch.bfloat16) if tail_mode != 1 else None
    tail = torch.full((R, 256), float("nan"), device=DEV, dtype=torch.bfloat16) if tail_mode == 1 else None
    native.check(native.load().ls_test_tblock(native.ptr(att), native.ptr(u), native.ptr(wo), native.ptr(w1),
                                              native.ptr(w2), native.ptr(wqkv), native.ptr(vec), native.ptr(qkv),
                                              native.ptr(tail), native.ptr(lengths), R, T or R, tail_mode,
                                              native.current_stream_ptr(DEV)), "ls_test_tblock")
    torch.cuda.synchronize()
    return qkv, tail


@pytest.mark.parametrize("R", [128, 300, 128 * 151 + 17])
def test_tblock_qkv_tail(R):
    ops = _tblock_operands(R, 5 + R)
    u_ref, qkv_ref = _tblock_ref(*ops)
    u = ops[1].clone()
    qkv, _ = _tblock(ops[0], u, *ops[2:], tail_mode=0)
    assert rel(u, u_ref) < 2e-3, rel(u, u_ref)       # bf16 rounding of the FF intermediate on both sides
    assert rel(qkv.float(), qkv_ref) < 6e-3, rel(qkv.float(), qkv_ref)


@pytest.mark.parametrize("R", [128, 300, 128 * 151 + 17])
def test_tblock_head_mode_layernorm_qkv_only(R):
    """tail_mode 2: qkv = LayerNorm(u; norm1) . Wqkv^T, u untouched (first block of a group)."""
    ops = _tblock_operands(R, 11 + R)
    att, u0, wo, w1, w2, wqkv, vec = ops
    g1n, be1n = vec[2048:2304], vec[2304:2560]
    n1 = bf16(F.layer_norm(u0, (256,), g1n, be1n, 1e-5)).float()
    qkv_ref = n1 @ wqkv.float().T
    u = u0.clone()
    qkv, _ = _tblock(att, u, wo, w1, w2, wqkv, vec, tail_mode=2)
    assert torch.equal(u, u0)
    assert rel(qkv.float(), qkv_ref) < 4e-3, rel(qkv.float(), qkv_ref)


def test_tblock_masked_copy_tail_and_tile_skip():
    T, lens = 200, [200, 70, 0, 131]
    R = T * len(lens)
    ops = _tblock_operands(R, 77)
    u_ref, _ = _tblock_ref(*ops)
    u = ops[1].clone()
    lengths = torch.tensor(lens, dtype=torch.int32, device=DEV)
    _, tail = _tblock(ops[0], u, *ops[2:], tail_mode=1, lengths=lengths, T=T)
    assert torch.equal(u, ops[1])  # tail_mode 1 leaves the residual stream alone
    tail = tail.float().view(len(lens), T, 256)
    u_ref = u_ref.view(len(lens), T, 256)
    for b, n in enumerate(lens):
        if n:
            assert rel(tail[b, :n], u_ref[b, :n]) < 4e-3
        # padded rows are zero, except inside tiles that are padding only (skipped: left untouched = NaN fill)
        pad = tail[b, n:]
        assert bool(((pad == 0) | pad.isnan()).all())
    rows = torch.arange(R, device=DEV).view(len(lens), T)
    for b, n in enumerate(lens):  # a padded row inside a tile that also holds valid rows must be exactly zero
        for t in range(n, T):
            r = int(rows[b, t])
            tile0 = (r // 128) * 128
            has_valid = any((rr // T) < len(lens) and (rr % T) < lens[rr // T] for rr in range(tile0, min(tile0 + 128, R)))
            if has_valid:
                assert float(tail[b, t].abs().max()) == 0.0
