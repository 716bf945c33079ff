// Bandwidth-bound kernels of the flow path: layout packing (the reference's einops.pack / rearrange copies,
// flow/decoder.py:427-433), timestep conditioning (matcha decoder.py:14-29,73-117 + ResnetBlock1D.mlp :49),
// CFG combine + Euler update (flow_matching.py:118-120) and the NCT <-> time-major boundary transposes.
#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

// ---- NCT fp32 -> time-major bf16 (32x32 smem transpose tiles) ----
__global__ void pack_nct_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int T,
                                long long src_bstride, int ld, int c_off, const int* __restrict__ lengths) {
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
    if (t < T && c < C) dst[((long long)b * T + t) * ld + c_off + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

__global__ void pack_bcast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int T,
                                  int ld, int c_off, const int* __restrict__ lengths) {
  const int b = blockIdx.y;
  const int len = lengths ? min(lengths[b], T) : T;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)T * C) return;
  const int t = (int)(idx / C), c = (int)(idx % C);
  dst[((long long)b * T + t) * ld + c_off + c] = __float2bfloat16(t < len ? src[(long long)b * C + c] : 0.f);
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
                                  int ld, const int* __restrict__ lengths) {
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
      const __nv_bfloat16 h = __float2bfloat16(v);
      xin[((long long)b * T + t) * ld + c] = h;
      xin[((long long)(B + b) * T + t) * ld + c] = h;
    }
  }
}

// v: [2B][T][C] fp32 time-major (conditional rows first).  4 channels per thread.
__global__ void cfg_euler_kernel(const float* __restrict__ v, float* __restrict__ x_state,
                                 __nv_bfloat16* __restrict__ xin, int B, int C, int T, int ld, float dt,
                                 float cfg_rate) {
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
    h.x = pack_bf16x2(x.x, x.y);
    h.y = pack_bf16x2(x.z, x.w);
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

__global__ void mask_to_lengths_kernel(const float* __restrict__ mask, int* __restrict__ lengths, int B, int T,
                                       int dup) {
  const int b = blockIdx.x;
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  int local = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) local += mask[(long long)b * T + t] != 0.f ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, local);
  __syncthreads();
  if (threadIdx.x == 0)
    for (int d = 0; d < dup; ++d) lengths[d * B + b] = cnt;
}

// ---- timestep conditioning: one block per time value ----
// out[o] = act_in(in)[:] . W[o][:] + bias[o], warp per output row, fp32
__device__ void block_gemv(const float* __restrict__ W, const float* __restrict__ bias, const float* in_smem,
                           float* out, int n_out, int n_in) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int o = warp; o < n_out; o += nwarps) {
    const float* w = W + (long long)o * n_in;
    float acc = 0.f;
    for (int k = lane; k < n_in; k += 32) acc = fmaf(__ldg(w + k), in_smem[k], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o] = acc + bias[o];
  }
}

__global__ void __launch_bounds__(1024) time_embed_kernel(const TimeEmbedParams p) {
  extern __shared__ float sm[];
  float* emb = sm;                 // [in_dim]
  float* h1 = emb + p.in_dim;      // [hid]
  float* h2 = h1 + p.hid;          // [hid]
  const int it = blockIdx.x;
  const float t = p.t[it];
  const int half = p.in_dim / 2;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = (1000.0f * t) * p.freqs[i];  // scale * x * emb  (matcha decoder.py:27)
    emb[i] = sinf(a);
    emb[half + i] = cosf(a);
  }
  __syncthreads();
  block_gemv(p.w1, p.b1, emb, h1, p.hid, p.in_dim);
  __syncthreads();
  for (int i = threadIdx.x; i < p.hid; i += blockDim.x) {
    const float x = h1[i];
    h1[i] = x / (1.0f + expf(-x));  // SiLU
  }
  __syncthreads();
  block_gemv(p.w2, p.b2, h1, h2, p.hid, p.hid);
  __syncthreads();
  for (int i = threadIdx.x; i < p.hid; i += blockDim.x) {  // Mish feeding every resnet's Linear
    const float x = h2[i];
    const float sp = x > 20.f ? x : log1pf(expf(x));
    h2[i] = x * tanhf(sp);
  }
  __syncthreads();
  for (int r = 0; r < p.n_res; ++r)
    block_gemv(p.wr + (long long)r * p.out_dim * p.hid, p.br + (long long)r * p.out_dim, h2,
               p.out + ((long long)it * p.n_res + r) * p.out_dim, p.out_dim, p.hid);
}

}  // namespace

cudaError_t launch_pack_nct(const float* src, __nv_bfloat16* dst, int B, int C, int T, long long src_bstride,
                            int ld, int c_off, const int* lengths, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  pack_nct_kernel<<<grid, block, 0, s>>>(src, dst, C, T, src_bstride, ld, c_off, lengths);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_pack_bcast(const float* src, __nv_bfloat16* dst, int B, int C, int T, int ld, int c_off,
                              const int* lengths, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  dim3 grid((unsigned)(((long long)T * C + 255) / 256), B);
  pack_bcast_kernel<<<grid, 256, 0, s>>>(src, dst, C, T, ld, c_off, lengths);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_pack_zero(__nv_bfloat16* dst, int B, int C, int T, int ld, int c_off, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  dim3 grid((unsigned)(((long long)T * C + 255) / 256), B);
  pack_zero_kernel<<<grid, 256, 0, s>>>(dst, C, T, ld, c_off);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_init_state(const float* noise, int noise_ld, float temperature, float* x_state,
                              __nv_bfloat16* xin, int B, int C, int T, int ld, const int* lengths, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  init_state_kernel<<<grid, block, 0, s>>>(noise, noise_ld, temperature, x_state, xin, B, C, T, ld, lengths);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_cfg_euler(const float* v, float* x_state, __nv_bfloat16* xin, int B, int C, int T, int ld,
                             float dt, float cfg_rate, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  const long long n4 = (long long)B * T * C / 4;
  int grid = (int)((n4 + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  if (grid < 1) grid = 1;
  cfg_euler_kernel<<<grid, 256, 0, s>>>(v, x_state, xin, B, C, T, ld, dt, cfg_rate);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_unpack_nct(const float* src, float* dst, int B, int C, int T, const int* lengths, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
  unpack_nct_kernel<<<grid, block, 0, s>>>(src, dst, C, T, lengths);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_mask_to_lengths(const float* mask, int* lengths, int B, int T, int dup, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  mask_to_lengths_kernel<<<B, 256, 0, s>>>(mask, lengths, B, T, dup);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_time_embed(const TimeEmbedParams& p, cudaStream_t s) {
  ProfScope prof(s, PK_ELEMENTWISE, 0.0, 0.0);
  if (p.nt <= 0) return cudaSuccess;
  const size_t smem = (size_t)(p.in_dim + 2 * p.hid) * sizeof(float);
  time_embed_kernel<<<p.nt, 1024, smem, s>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ls
