// Bandwidth-bound kernels of the flow path: layout packing (the reference's einops.pack / rearrange copies,
// flow/decoder.py:427-433), timestep conditioning (matcha decoder.py:14-29,73-117 + ResnetBlock1D.mlp :49),
// CFG combine + Euler update (flow_matching.py:118-120) and the NCT <-> time-major boundary transposes.
#include <algorithm>

#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

// 16-bit operand format of the destination: bf16 (default) or fp16 (the flow estimator's fp16-operand mode)
__device__ __forceinline__ void store_h(__nv_bfloat16* dst, long long i, float v, int fp16) {
  reinterpret_cast<uint16_t*>(dst)[i] = fp16 ? cvt_f16_bits(v) : cvt_bf16_bits(v);
}

// ---- NCT fp32 -> time-major bf16 (32x32 smem transpose tiles) ----
__global__ void pack_nct_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int T,
                                long long src_bstride, int ld, int c_off, const int* __restrict__ lengths, int fp16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = lengths ? min(lengths[b], T) : T;
  const float* s = src + (long long)b * src_bstride;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < len) ? s[(long long)c * T + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) store_h(dst, ((long long)b * T + t) * ld + c_off + c, tile[threadIdx.x][i], fp16);
  }
}

__global__ void pack_bcast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int T,
                                  int ld, int c_off, const int* __restrict__ lengths, int fp16) {
  const int b = blockIdx.y;
  const int len = lengths ? min(lengths[b], T) : T;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * C) return;
  const int t = (int)(idx / C), c = (int)(idx % C);
  store_h(dst, ((long long)b * T + t) * ld + c_off + c, t < len ? src[(long long)b * C + c] : 0.f, fp16);
}

__global__ void pack_zero_kernel(__nv_bfloat16* __restrict__ dst, int C, int T, int ld, int c_off) {
  const int b = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * C) return;
  const int t = (int)(idx / C), c = (int)(idx % C);
  dst[((long long)b * T + t) * ld + c_off + c] = __float2bfloat16(0.f);
}

// x_state[b][t][c] = noise[c][t]*temperature ; bf16 copies into xin rows b and B+b
__global__ void init_state_kernel(const float* __restrict__ noise, int noise_ld, float temperature,
                                  float* __restrict__ x_state, __nv_bfloat16* __restrict__ xin, int B, int C, int T,
                                  int ld, const int* __restrict__ lengths, int fp16) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = lengths ? min(lengths[b], T) : T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && t < len) ? noise[(long long)c * noise_ld + t] * temperature : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    if (t < T && c < C) {
      const float v = tile[threadIdx.x][i];
      x_state[((long long)b * T + t) * C + c] = v;
      store_h(xin, ((long long)b * T + t) * ld + c, v, fp16);
      store_h(xin, ((long long)(B + b) * T + t) * ld + c, v, fp16);
    }
  }
}

// v: [2B][T][C] fp32 time-major (conditional rows first).  4 channels per thread.
__global__ void cfg_euler_kernel(const float* __restrict__ v, float* __restrict__ x_state,
                                 __nv_bfloat16* __restrict__ xin, int B, int C, int T, int ld, float dt,
                                 float cfg_rate, int fp16) {
  const long long n4 = (long long)B * T * C / 4;
  const long long half = (long long)B * T * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 vc = __ldg(reinterpret_cast<const float4*>(v) + i);
    const float4 vu = __ldg(reinterpret_cast<const float4*>(v + half) + i);
    float4 x = reinterpret_cast<float4*>(x_state)[i];
    // reference order: dphi = (1+w)*vc - w*vu ; x = x + dt*dphi
    x.x = x.x + dt * ((1.0f + cfg_rate) * vc.x - cfg_rate * vu.x);
    x.y = x.y + dt * ((1.0f + cfg_rate) * vc.y - cfg_rate * vu.y);
    x.z = x.z + dt * ((1.0f + cfg_rate) * vc.z - cfg_rate * vu.z);
    x.w = x.w + dt * ((1.0f + cfg_rate) * vc.w - cfg_rate * vu.w);
    reinterpret_cast<float4*>(x_state)[i] = x;
    const long long e = i * 4;
    const long long row = e / C;  // b*T + t
    const int c = (int)(e % C);
    uint2 h;
    h.x = fp16 ? pack_f16x2(x.x, x.y) : pack_bf16x2(x.x, x.y);
    h.y = fp16 ? pack_f16x2(x.z, x.w) : pack_bf16x2(x.z, x.w);
    *reinterpret_cast<uint2*>(xin + row * ld + c) = h;
    *reinterpret_cast<uint2*>(xin + (row + (long long)B * T) * ld + c) = h;
  }
}

__global__ void unpack_nct_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int T,
                                  const int* __restrict__ lengths) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = lengths ? min(lengths[b], T) : T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < len && c < C) ? src[((long long)b * T + t) * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) dst[((long long)b * C + c) * T + t] = tile[threadIdx.x][i];
  }
}

// lengths[b] = number of non-zero mask entries; a mask that is not a prefix (right-padding) mask -- a non-zero entry after
// a zero -- raises *bad_flag (optional: a sticky flag the engine reports at its next call, so that validating the mask
// costs no host synchronisation on the hot call)
__global__ void mask_to_lengths_kernel(const float* __restrict__ mask, int* __restrict__ lengths, int B, int T,
                                       int dup, int* bad_flag) {
  const int b = blockIdx.x;
  __shared__ int cnt, bad;
  if (threadIdx.x == 0) cnt = 0, bad = 0;
  __syncthreads();
  int local = 0, local_bad = 0;
  const float* m = mask + (long long)b * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const bool on = m[t] != 0.f;
    local += on ? 1 : 0;
    if (on && t > 0 && m[t - 1] == 0.f) local_bad = 1;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  local_bad = __any_sync(0xffffffffu, local_bad);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&cnt, local);
    if (local_bad) bad = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int d = 0; d < dup; ++d) lengths[d * B + b] = cnt;
    if (bad && bad_flag) *bad_flag = 1;
  }
}

// ---- GroupNorm(8) + Mish of the non-causal ConditionalDecoder's blocks (matcha Block1D, decoder.py:32-43) on time-major
// fp32 [B][T][C].  The statistics of a (batch row, group) run over (C/G channels x the row's valid frames), so they cannot
// ride in the producing GEMM's epilogue (a tile sees 128 frames): one pass for the statistics, one to apply them.
// Stats: one block per (group, batch row); per-thread Welford over its rows, merged with Chan's formula.
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x, float2* __restrict__ stats, int C,
                                                              int T, int G, const int* __restrict__ lengths) {
  const int grp = blockIdx.x, b = blockIdx.y;
  const int cpg = C / G;  // 32: one warp reads one frame's channels of the group (128 B)
  const int len = lengths ? min(lengths[b], T) : T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int t = warp; t < len; t += nw)
    for (int c = lane; c < cpg; c += 32) {
      const float v = x[((long long)b * T + t) * C + grp * cpg + c];
      n += 1.f;
      const float d = v - mean;
      mean += d / n;
      m2 = fmaf(d, v - mean, m2);
    }
  auto merge = [](float& na, float& ma, float& qa, float nb, float mb, float qb) {
    const float nn = na + nb;
    if (nn > 0.f) {
      const float d = mb - ma;
      ma += d * (nb / nn);
      qa += qb + d * d * (na * nb / nn);
      na = nn;
    }
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o), mb = __shfl_xor_sync(0xffffffffu, mean, o),
                qb = __shfl_xor_sync(0xffffffffu, m2, o);
    merge(n, mean, m2, nb, mb, qb);
  }
  __shared__ float sn[8], sm[8], sq[8];
  if (lane == 0) sn[warp] = n, sm[warp] = mean, sq[warp] = m2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float na = sn[0], ma = sm[0], qa = sq[0];
    for (int w = 1; w < nw; ++w) merge(na, ma, qa, sn[w], sm[w], sq[w]);
    stats[b * G + grp] = make_float2(ma, na > 0.f ? rsqrtf(qa / na + 1e-5f) : 0.f);
  }
}
// y = mish((x - mean) rstd gamma + beta) (+ temb[b][c]) (+ addend fp32) on valid frames, 0 on padding (what the next
// convolution's x * mask makes of them); written as fp32 and / or in the 16-bit operand format.  4 channels per thread.
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, const float2* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ temb, long long temb_bstride,
                                                              const float* __restrict__ addend, float* __restrict__ out_f32,
                                                              __nv_bfloat16* __restrict__ out_h, int B, int C, int T, int G,
                                                              const int* __restrict__ lengths, int fp16) {
  const long long n4 = (long long)B * T * C / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const long long row = e / C;
    const int c = (int)(e - row * C);
    const int b = (int)(row / T), t = (int)(row - (long long)b * T);
    const int len = lengths ? min(lengths[b], T) : T;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len) {
      const float4 v = reinterpret_cast<const float4*>(x)[i];
      const float2 st = stats[b * G + c / (C / G)];
      const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
      y.x = mish_f(fmaf((v.x - st.x) * st.y, gm.x, bt.x));
      y.y = mish_f(fmaf((v.y - st.x) * st.y, gm.y, bt.y));
      y.z = mish_f(fmaf((v.z - st.x) * st.y, gm.z, bt.z));
      y.w = mish_f(fmaf((v.w - st.x) * st.y, gm.w, bt.w));
      if (temb) {
        const float4 tv = *reinterpret_cast<const float4*>(temb + (long long)b * temb_bstride + c);
        y.x += tv.x, y.y += tv.y, y.z += tv.z, y.w += tv.w;
      }
      if (addend) {
        const float4 a = reinterpret_cast<const float4*>(addend)[i];
        y.x += a.x, y.y += a.y, y.z += a.z, y.w += a.w;
      }
    }
    if (out_f32) reinterpret_cast<float4*>(out_f32)[i] = y;
    if (out_h) {
      uint2 h;
      h.x = fp16 ? pack_f16x2(y.x, y.y) : pack_bf16x2(y.x, y.y);
      h.y = fp16 ? pack_f16x2(y.z, y.w) : pack_bf16x2(y.z, y.w);
      reinterpret_cast<uint2*>(out_h)[i] = h;
    }
  }
}

// ---- FSQ quantizer head of the S3 speech tokenizer (tools/S3Tokenizer/s3tokenizer/model_v2.py:83-117, FSQCodebook.encode):
// h = round(tanh(project_down(x)) * 0.999) + 1 in {0, 1, 2}^8, token = sum_i h_i 3^i.  One warp per frame: eight fp32 dot
// products over the hidden dimension, torch.round semantics (half to even).
__global__ void __launch_bounds__(256) fsq_encode_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int* __restrict__ tokens,
                                                         long long rows, int D) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * D;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = xr[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, __ldg(w + (long long)j * D + k), acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane == 0) {
    int tok = 0, p3 = 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float h = tanhf(acc[j] + bias[j]) * 0.9990000128746033f;
      tok += ((int)rintf(h) + 1) * p3;
      p3 *= 3;
    }
    tokens[row] = tok;
  }
}

// ---- timestep conditioning (matcha decoder.py:14-29 SinusoidalPosEmb, :73-117 TimestepEmbedding, :49 ResnetBlock1D.mlp)
// Three GEMV stages over ALL time values of a solve at once, each a grid of (output rows / 8, nt) blocks of 8 warps with
// one output row per warp -- the weights (19 MB fp32) are streamed by ~450 blocks per time value instead of by one
// 1024-thread block per time value (634 us per solve before, launch-latency-sized now).
//   stage 0: in = [sin(1000 t f_i) | cos(1000 t f_i)]         out = W1 in + b1
//   stage 1: in = SiLU(prev)                                  out = W2 in + b2
//   stage 2: in = Mish(prev)                                  out[r] = Wr[r] in + br[r]   (every resnet's Linear, stacked)
constexpr int kTimeInline = 64;  // time values carried in the kernel parameters (keeps the solve capturable in a CUDA graph)
struct TimeStage {
  const float* t;         // [nt] device, or nullptr: t_inline
  float t_inline[kTimeInline];
  const float* freqs;     // stage 0
  const float* in;        // stages 1, 2: [nt][n_in]
  const float* W;         // [n_out][n_in]
  const float* bias;      // [n_out]
  float* out;             // [nt][n_out]
  int n_in, n_out;
};
template <int kStage>
__global__ void __launch_bounds__(256) time_stage_kernel(const TimeStage p) {
  extern __shared__ float vin[];  // [n_in]
  const int it = blockIdx.y;
  if (kStage == 0) {
    const float t = p.t ? p.t[it] : p.t_inline[it];
    const int half = p.n_in / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
      const float a = (1000.0f * t) * p.freqs[i];  // scale * x * emb  (matcha decoder.py:27)
      vin[i] = sinf(a);
      vin[half + i] = cosf(a);
    }
  } else {
    const float* src = p.in + (long long)it * p.n_in;
    for (int i = threadIdx.x; i < p.n_in; i += blockDim.x) {
      const float x = src[i];
      if (kStage == 1) {
        vin[i] = x / (1.0f + expf(-x));  // SiLU
      } else {
        const float sp = x > 20.f ? x : log1pf(expf(x));  // Mish feeding every resnet's Linear
        vin[i] = x * tanhf(sp);
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  if (o >= p.n_out) return;
  const float* w = p.W + (long long)o * p.n_in;
  float acc = 0.f;
  for (int k = lane; k < p.n_in; k += 32) acc = fmaf(__ldg(w + k), vin[k], acc);
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) p.out[(long long)it * p.n_out + o] = acc + p.bias[o];
}

}  // namespace

cudaError_t launch_pack_nct(const float* src, __nv_bfloat16* dst, int B, int C, int T, long long src_bstride,
                            int ld, int c_off, const int* lengths, cudaStream_t s, int fp16) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * C * T * (4.0 + 2.0));
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  pack_nct_kernel<<<grid, block, 0, s>>>(src, dst, C, T, src_bstride, ld, c_off, lengths, fp16);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_pack_bcast(const float* src, __nv_bfloat16* dst, int B, int C, int T, int ld, int c_off,
                              const int* lengths, cudaStream_t s, int fp16) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * C * 4.0 + (double)B * C * T * 2.0);
  dim3 grid((unsigned)(((long long)T * C + 255) / 256), B);
  pack_bcast_kernel<<<grid, 256, 0, s>>>(src, dst, C, T, ld, c_off, lengths, fp16);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_pack_zero(__nv_bfloat16* dst, int B, int C, int T, int ld, int c_off, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * C * T * 2.0);
  dim3 grid((unsigned)(((long long)T * C + 255) / 256), B);
  pack_zero_kernel<<<grid, 256, 0, s>>>(dst, C, T, ld, c_off);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_init_state(const float* noise, int noise_ld, float temperature, float* x_state,
                              __nv_bfloat16* xin, int B, int C, int T, int ld, const int* lengths, cudaStream_t s,
                              int fp16) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)C * T * 4.0 + (double)B * C * T * (4.0 + 2.0 * 2.0));
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  init_state_kernel<<<grid, block, 0, s>>>(noise, noise_ld, temperature, x_state, xin, B, C, T, ld, lengths, fp16);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_cfg_euler(const float* v, float* x_state, __nv_bfloat16* xin, int B, int C, int T, int ld,
                             float dt, float cfg_rate, cudaStream_t s, int fp16) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * C * T * (2 * 4.0 + 4.0 + 4.0 + 2 * 2.0));
  const long long n4 = (long long)B * T * C / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  cfg_euler_kernel<<<grid, 256, 0, s>>>(v, x_state, xin, B, C, T, ld, dt, cfg_rate, fp16);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_unpack_nct(const float* src, float* dst, int B, int C, int T, const int* lengths, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * C * T * (4.0 + 4.0));
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  unpack_nct_kernel<<<grid, block, 0, s>>>(src, dst, C, T, lengths);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_mask_to_lengths(const float* mask, int* lengths, int B, int T, int dup, cudaStream_t s, int* bad_flag) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, (double)B * T * 4.0);
  mask_to_lengths_kernel<<<B, 256, 0, s>>>(mask, lengths, B, T, dup, bad_flag);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_fsq_encode(const float* x, const float* w, const float* bias, int* tokens, long long rows, int D,
                              cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  ProfScope prof(s, PK_ELEMENTWISE, 16.0 * rows * D, (double)rows * D * 4.0 + 8.0 * D * 4.0 + rows * 4.0);
  fsq_encode_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, w, bias, tokens, rows, D);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_groupnorm_mish(const float* x, float2* stats, const float* gamma, const float* beta, const float* temb,
                                  long long temb_bstride, const float* addend, float* out_f32, __nv_bfloat16* out_h, int B,
                                  int C, int T, int G, const int* lengths, int fp16, cudaStream_t s) {
  if (C % (4 * G) || (C / G) % 32) return cudaErrorInvalidValue;
  const double elems = (double)B * T * C;
  {
    ProfScope prof(s, PK_ELEMENTWISE, 0.0, elems * 4.0);
    groupnorm_stats_kernel<<<dim3(G, B), 256, 0, s>>>(x, stats, C, T, G, lengths);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, elems * (4.0 + (addend ? 4.0 : 0.0) + (out_f32 ? 4.0 : 0.0) + (out_h ? 2.0 : 0.0)));
  const long long n4 = (long long)B * T * C / 4;
  int grid = (int)std::min<long long>((n4 + 255) / 256, 148 * 8);
  groupnorm_apply_kernel<<<grid < 1 ? 1 : grid, 256, 0, s>>>(x, stats, gamma, beta, temb, temb_bstride, addend, out_f32, out_h, B,
                                                            C, T, G, lengths, fp16);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_time_embed(const TimeEmbedParams& p, cudaStream_t s) {
  if (p.nt <= 0) return cudaSuccess;
  if (p.t == nullptr && p.nt > kTimeInline) return cudaErrorInvalidValue;
  const double w_bytes = 4.0 * ((double)p.hid * p.in_dim + (double)p.hid * p.hid + (double)p.n_res * p.out_dim * p.hid);
  TimeStage st{};
  st.t = p.t;
  if (!p.t)
    for (int i = 0; i < p.nt; ++i) st.t_inline[i] = p.t_host[i];
  auto run = [&](auto kernel, int n_in, int n_out, double flops, double bytes) -> cudaError_t {
    ProfScope prof(s, PK_ELEMENTWISE, flops, bytes);
    st.n_in = n_in, st.n_out = n_out;
    kernel<<<dim3((n_out + 7) / 8, p.nt), 256, (size_t)n_in * sizeof(float), s>>>(st);
    count_launch();
    return cudaGetLastError();
  };
  (void)w_bytes;
  // h1 [nt][hid] and h2 [nt][hid] live behind the output (the caller's buffer has room: see TimeEmbedParams::scratch)
  st.freqs = p.freqs, st.W = p.w1, st.bias = p.b1, st.out = p.scratch;
  cudaError_t e = run(time_stage_kernel<0>, p.in_dim, p.hid, 2.0 * p.nt * p.hid * p.in_dim, 4.0 * p.hid * (p.in_dim + 1.0 + p.nt));
  if (e != cudaSuccess) return e;
  st.in = p.scratch, st.W = p.w2, st.bias = p.b2, st.out = p.scratch + (long long)p.nt * p.hid;
  e = run(time_stage_kernel<1>, p.hid, p.hid, 2.0 * p.nt * p.hid * p.hid, 4.0 * p.hid * (p.hid + 1.0 + 2.0 * p.nt));
  if (e != cudaSuccess) return e;
  st.in = p.scratch + (long long)p.nt * p.hid, st.W = p.wr, st.bias = p.br, st.out = p.out;
  const int n_out = p.n_res * p.out_dim;
  return run(time_stage_kernel<2>, p.hid, n_out, 2.0 * p.nt * n_out * p.hid, 4.0 * n_out * (p.hid + 1.0 + p.nt) + 4.0 * p.nt * p.hid);
}

}  // namespace ls
