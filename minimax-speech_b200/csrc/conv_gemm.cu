// Implicit-GEMM conv1d / linear kernel for sm_100a.
//
// Replaces, on time-major bf16 activations, every dense contraction of the hot path:
//   CausalConv1d k=3          speech/cosyvoice/flow/decoder.py:36-62
//   res_conv / final_proj 1x1 speech/matcha/models/components/decoder.py:54,60 ; flow/decoder.py:403
//   to_q/k/v, to_out, FF      speech/matcha/models/components/transformer.py:196-204,109-126
//   WNConv1d k=7 dilated, k=1 dac-vae/model.py:128-130,343,364
//   WNConvTranspose1d         dac-vae/model.py:255-262 (two-tap polyphase form, N = stride*Cout)
// with the surrounding row-local work fused into the epilogue (bias, LeakyReLU / exact GELU / LayerNorm+Mish /
// tanh, time-embedding add, residual add, length masking, and a second bf16 output holding LayerNorm / Snake /
// copy of the result for the next GEMM).
//
// Structure: persistent CTAs (one per SM), warp-specialised:
//   warp 0  TMA producer  (A: 128 rows x 64 ch box per tap via a 3-D map, OOB rows -> 0 = conv zero padding;
//                          B: block_n x 64 weight box), multi-stage mbarrier ring
//   warp 1  tcgen05.mma issuer (one lane), fp32 accumulators in TMEM, double buffered
//   warp 2  TMEM allocator
//   warps 4-7 epilogue: thread = output row, tcgen05.ld 16 columns at a time
#include "kernels.h"
#include "ptx.cuh"

namespace ls {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kThreads = 256;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kMaxStages = 6;
constexpr int kSmemBudget = 220 * 1024;

struct TileCoord {
  int b, mt, nt;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, int m_tiles, int n_tiles) {
  TileCoord c;
  c.nt = tile % n_tiles;
  const int rest = tile / n_tiles;
  c.mt = rest % m_tiles;
  c.b = rest / m_tiles;
  return c;
}

__device__ __forceinline__ bool tile_skipped(const ConvGemmParams& p, const TileCoord& c) {
  if (p.lengths == nullptr) return false;
  const long long need = (long long)p.lengths[c.b] * p.m_len_mul + p.m_len_add + p.skip_halo;
  return (long long)c.mt * kBlockM >= need;
}

// Stores 16 consecutive values u[0..15] of this thread's row starting at column n (n % 16 == 0).
// flat = row_flat + n is the element index inside the batch item.
struct RowStore {
  long long valid, alloc;
  int n_store;
  bool row_in;

  __device__ __forceinline__ int group_state(long long flat, int len) const {  // 2 store, 1 zero, 0 drop
    if (!row_in || flat < 0) return 0;
    if (flat + len <= valid) return 2;
    if (flat + len <= alloc) return 1;
    return 0;
  }
  __device__ __forceinline__ void f32(float* base, long long flat, int n, const float (&u)[16]) const {
    if (n_store - n >= 16 || ((n_store - n) & 3) == 0) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (n + 4 * g >= n_store) break;
        const int st = group_state(flat + 4 * g, 4);
        if (st == 0) continue;
        float4 v = st == 2 ? make_float4(u[4 * g], u[4 * g + 1], u[4 * g + 2], u[4 * g + 3])
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(base + flat + 4 * g) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (n + i >= n_store) break;
        const int st = group_state(flat + i, 1);
        if (st) base[flat + i] = st == 2 ? u[i] : 0.f;
      }
    }
  }
  __device__ __forceinline__ void bf16(__nv_bfloat16* base, long long flat, int n, const float (&u)[16]) const {
    if (n_store - n >= 16 || ((n_store - n) & 7) == 0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (n + 8 * g >= n_store) break;
        const int st = group_state(flat + 8 * g, 8);
        if (st == 0) continue;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (st == 2) {
          v.x = pack_bf16x2(u[8 * g + 0], u[8 * g + 1]);
          v.y = pack_bf16x2(u[8 * g + 2], u[8 * g + 3]);
          v.z = pack_bf16x2(u[8 * g + 4], u[8 * g + 5]);
          v.w = pack_bf16x2(u[8 * g + 6], u[8 * g + 7]);
        }
        *reinterpret_cast<uint4*>(base + flat + 8 * g) = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (n + i >= n_store) break;
        const int st = group_state(flat + i, 1);
        if (st) base[flat + i] = __float2bfloat16(st == 2 ? u[i] : 0.f);
      }
    }
  }
};

__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + g);
    v[4 * g] = t.x, v[4 * g + 1] = t.y, v[4 * g + 2] = t.z, v[4 * g + 3] = t.w;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapW, const __grid_constant__ ConvGemmParams p,
                 const int stages, const int tmem_cols, const int acc_stride) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = kABytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = p.N / p.block_n;
  const int total_tiles = p.B * m_tiles * n_tiles;
  const int k_iters = p.taps * p.kb_per_tap;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA0);
    prefetch_tmap(&mapA1);
    prefetch_tmap(&mapW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, m_tiles, n_tiles);
        if (tile_skipped(p, tc)) continue;
        const int t0 = tc.mt * kBlockM;
        const int n0 = tc.nt * p.block_n;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int trow = t0 + tap * p.dil - p.pad;
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            mbar_arrive_expect_tx(&full[stage], (uint32_t)stage_bytes);
            if (kb < p.kb_split)
              tma_load_3d(sa, &mapA0, &full[stage], kb * kBlockK, trow, tc.b);
            else
              tma_load_3d(sa, &mapA1, &full[stage], (kb - p.kb_split) * kBlockK, trow, tc.b);
            tma_load_2d(sa + kABytes, &mapW, &full[stage], kb * kBlockK, tap * p.N + n0);
            if (++stage == stages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(tile, m_tiles, n_tiles);
        if (tile_skipped(p, tc)) continue;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == stages) stage = 0, phase ^= 1;
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (thread = output row)
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, m_tiles, n_tiles);
      if (tile_skipped(p, tc)) continue;
      const int t = tc.mt * kBlockM + row;
      const int n0 = tc.nt * p.block_n;
      RowStore st;
      st.alloc = p.out_alloc;
      st.valid = p.lengths ? min((long long)p.lengths[tc.b] * p.out_valid_mul, p.out_alloc) : p.out_alloc;
      st.n_store = p.n_store;
      st.row_in = t < p.M;
      const long long row_flat = (long long)t * p.out_ld + p.out_shift;
      const long long boff = (long long)tc.b * p.out_bstride;
      float* out0f = reinterpret_cast<float*>(p.out0) + boff;
      __nv_bfloat16* out0h = reinterpret_cast<__nv_bfloat16*>(p.out0) + boff;
      __nv_bfloat16* out1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + boff;
      const float* addf = reinterpret_cast<const float*>(p.addend) + boff;
      const __nv_bfloat16* addh = reinterpret_cast<const __nv_bfloat16*>(p.addend) + boff;
      const float* temb = p.temb ? p.temb + (long long)tc.b * p.temb_bstride : nullptr;

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * acc_stride);

      // v = acc + bias, then the cheap activations
      auto pre = [&](int c, float (&v)[16]) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c, r);
        tmem_ld_wait();
        const int ch0 = (n0 + c) % p.chan_mod;
        if (p.bias) {
          float bv[16];
          load16(p.bias + ch0, bv);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bv[i];
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (p.act == ACT_LRELU || p.act == ACT_LRELU_TANH) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.1f * v[i];
          if (p.act == ACT_LRELU_TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
          }
        } else if (p.act == ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
        }
      };

      float mean = 0.f, rstd = 0.f;
      if (p.act == ACT_LN_MISH) {  // LayerNorm statistics over the full row (block_n == N)
        float s = 0.f, ss = 0.f, shift = 0.f;
        for (int c = 0; c < p.block_n; c += 16) {
          float v[16];
          pre(c, v);
          if (c == 0) shift = v[0];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float d = v[i] - shift;
            s += d;
            ss = fmaf(d, d, ss);
          }
        }
        const float inv_n = 1.0f / (float)p.block_n;
        const float md = s * inv_n;
        mean = shift + md;
        rstd = rsqrtf(fmaxf(ss * inv_n - md * md, 0.f) + 1e-5f);
      }

      float s1 = 0.f, ss1 = 0.f, shift1 = 0.f;  // statistics of u for OUT1_LN
      for (int c = 0; c < p.block_n; c += 16) {
        float u[16];
        pre(c, u);
        const int n = n0 + c;
        const int ch0 = n % p.chan_mod;
        if (p.act == ACT_LN_MISH) {
          float g[16], bb[16];
          load16(p.ln_g + ch0, g);
          load16(p.ln_b + ch0, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = mish_f(fmaf((u[i] - mean) * rstd, g[i], bb[i]));
        }
        if (temb) {
          float tv[16];
          load16(temb + n, tv);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] += tv[i];
        }
        const long long flat = row_flat + n;
        if (p.addend && st.row_in && flat >= 0 && flat + 16 <= st.valid) {
          if (p.addend_dtype == OUT_F32) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float4 a = *reinterpret_cast<const float4*>(addf + flat + 4 * g4);
              u[4 * g4] += a.x, u[4 * g4 + 1] += a.y, u[4 * g4 + 2] += a.z, u[4 * g4 + 3] += a.w;
            }
          } else {
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              const uint4 a = *reinterpret_cast<const uint4*>(addh + flat + 8 * g8);
              const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
                u[8 * g8 + 2 * j] += __low2float(h2);
                u[8 * g8 + 2 * j + 1] += __high2float(h2);
              }
            }
          }
        }
        if (p.out0_dtype == OUT_F32)
          st.f32(out0f, flat, n, u);
        else if (p.out0_dtype == OUT_BF16)
          st.bf16(out0h, flat, n, u);
        if (p.out1_mode == OUT1_LN) {
          if (c == 0) shift1 = u[0];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float d = u[i] - shift1;
            s1 += d;
            ss1 = fmaf(d, d, ss1);
          }
        } else if (p.out1_mode == OUT1_COPY) {
          st.bf16(out1, flat, n, u);
        } else if (p.out1_mode == OUT1_SNAKE) {
          float al[16], ia[16];
          load16(p.p1_a + ch0, al);
          load16(p.p1_b + ch0, ia);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = snake_f(u[i], al[i], ia[i]);
          st.bf16(out1, flat, n, u);
        }
      }
      // TMEM accumulator fully consumed: hand it back to the MMA warp before the LN write-out
      tc_fence_before();
      mbar_arrive(&tempty[acc]);

      if (p.out1_mode == OUT1_LN) {  // second output = LayerNorm(u) in bf16; u re-read from this thread's own row
        const float inv_n = 1.0f / (float)p.block_n;
        const float md = s1 * inv_n;
        const float mean1 = shift1 + md;
        const float rstd1 = rsqrtf(fmaxf(ss1 * inv_n - md * md, 0.f) + 1e-5f);
        for (int c = 0; c < p.block_n; c += 16) {
          const int n = n0 + c;
          const long long flat = row_flat + n;
          float u[16], g[16], bb[16];
          if (st.row_in && flat >= 0 && flat + 16 <= st.valid) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float4 a = *reinterpret_cast<const float4*>(out0f + flat + 4 * g4);
              u[4 * g4] = a.x, u[4 * g4 + 1] = a.y, u[4 * g4 + 2] = a.z, u[4 * g4 + 3] = a.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) u[i] = 0.f;
          }
          load16(p.p1_a + n % p.chan_mod, g);
          load16(p.p1_b + n % p.chan_mod, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = fmaf((u[i] - mean1) * rstd1, g[i], bb[i]);
          st.bf16(out1, flat, n, u);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
  }
}

int pow2_at_least(int x, int lo) {
  int v = lo;
  while (v < x) v <<= 1;
  return v;
}

}  // namespace

cudaError_t launch_conv_gemm(const CUtensorMap& mapA0, const CUtensorMap& mapA1, const CUtensorMap& mapW,
                             const ConvGemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.block_n % 16 || p.block_n < 16 || p.block_n > 256 || p.N % p.block_n || p.chan_mod % 16) return cudaErrorInvalidValue;
  if ((p.act == ACT_LN_MISH || p.out1_mode == OUT1_LN) && p.block_n != p.N) return cudaErrorInvalidValue;
  if (p.out1_mode == OUT1_LN && p.out0_dtype != OUT_F32) return cudaErrorInvalidValue;
  if (p.out1_mode != OUT1_NONE && p.out1 == nullptr) return cudaErrorInvalidValue;
  const int stage_bytes = kABytes + p.block_n * kBlockK * 2;
  int stages = kSmemBudget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return cudaErrorInvalidValue;
  const int acc_stride = pow2_at_least(p.block_n, 32);
  const int tmem_cols = 2 * acc_stride;
  size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  // tmem_cols == 512 must never share an SM with a second CTA of this kernel (alloc would spin):
  if (tmem_cols > 256 && smem < 120 * 1024) smem = 120 * 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const long long total = (long long)p.B * m_tiles * (p.N / p.block_n);
  if (total <= 0) return cudaSuccess;
  const int grid = (int)(total < num_sms ? total : num_sms);
  conv_gemm_kernel<<<grid, kThreads, smem, stream>>>(mapA0, mapA1, mapW, p, stages, tmem_cols, acc_stride);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ls
