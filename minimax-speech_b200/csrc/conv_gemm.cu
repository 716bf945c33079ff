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
//   warp 0  TMEM allocator, then TMA producer  (A: 128 rows x 64 ch box per tap via a 3-D map, OOB rows -> 0 = conv zero padding;
//                          B: block_n x 64 weight box), multi-stage mbarrier ring
//   warp 1  tcgen05.mma issuer (one lane), fp32 accumulators in TMEM, double buffered
//   warps 2-17 epilogue (512 threads): warp w owns TMEM lanes 32*(w%4).. (= 32 output rows) and every 4th
//              16-column chunk; accumulators are pulled into registers in one burst and the TMEM stage is
//              released immediately, LayerNorm statistics are combined across the 4 column groups through smem
#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kFirstEpiWarp = 2;
constexpr int kThreads = kFirstEpiWarp * 32 + kEpiThreads;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kMaxStages = 6;
constexpr int kRedBytes = 2 * 2 * 4 * kBlockM * 8;  // [tile parity][phase][column group][row] float2
constexpr int kSmemBudget = 206 * 1024;             // for the operand ring

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

// 16 consecutive output columns of one row.  state: 2 store values, 1 store zeros, 0 drop.
struct ChunkStore {
  long long valid, alloc;
  int n_store;
  bool row_in;
  __device__ __forceinline__ int state(long long flat, int width = 16) const {
    if (!row_in || flat < 0) return 0;
    if (flat + width <= valid) return 2;
    if (flat + width <= alloc) return 1;
    return 0;
  }
  __device__ __forceinline__ void f32(float* base, long long flat, int n, const float (&u)[16]) const {
    const int st = state(flat, min(16, n_store - n));
    if (st == 0 || n >= n_store) return;
    if (n + 16 <= n_store) {
      float4* d = reinterpret_cast<float4*>(base + flat);
#pragma unroll
      for (int g = 0; g < 4; ++g)
        d[g] = st == 2 ? make_float4(u[4 * g], u[4 * g + 1], u[4 * g + 2], u[4 * g + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n + i < n_store) base[flat + i] = st == 2 ? u[i] : 0.f;
    }
  }
  __device__ __forceinline__ void bf16(__nv_bfloat16* base, long long flat, int n, const float (&u)[16]) const {
    const int st = state(flat, min(16, n_store - n));
    if (st == 0 || n >= n_store) return;
    if (n + 16 <= n_store) {
      uint4* d = reinterpret_cast<uint4*>(base + flat);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (st == 2) {
          v.x = pack_bf16x2(u[8 * g + 0], u[8 * g + 1]);
          v.y = pack_bf16x2(u[8 * g + 2], u[8 * g + 3]);
          v.z = pack_bf16x2(u[8 * g + 4], u[8 * g + 5]);
          v.w = pack_bf16x2(u[8 * g + 6], u[8 * g + 7]);
        }
        d[g] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n + i < n_store) base[flat + i] = __float2bfloat16(st == 2 ? u[i] : 0.f);
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

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// LayerNorm statistics of a 256-wide row whose 4 x 64 column quarters live in 4 threads (one per column
// group): per-thread (mean, M2), combined with the parallel-variance formula through shared memory.
__device__ __forceinline__ void row_stats(const float (&v)[4][16], float2* red, int g, int row, float& mean,
                                          float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[j][i];
  const float lm = s * (1.0f / 64.0f);
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[j][i] - lm;
      m2 = fmaf(d, d, m2);
    }
  red[g * kBlockM + row] = make_float2(lm, m2);
  epi_barrier();
  const float2 a = red[row], b = red[kBlockM + row], c = red[2 * kBlockM + row], d = red[3 * kBlockM + row];
  mean = 0.25f * (a.x + b.x + c.x + d.x);
  const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
  const float M2 = a.y + b.y + c.y + d.y + 64.0f * (da * da + db * db + dc * dc + dd * dd);
  rstd = rsqrtf(M2 * (1.0f / 256.0f) + 1e-5f);
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
  float2* red_base = reinterpret_cast<float2*>(bars + 32);  // 256 B after the barriers

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
      mbar_init(&tempty[i], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
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
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;                     // TMEM lane quarter this warp may read (warp id % 4)
    const int g = (warp - kFirstEpiWarp) >> 2;  // column group: owns 16-column chunks g, g+4, g+8, g+12
    const int row = q * 32 + lane;
    const int n_chunks = p.block_n >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t parity = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, m_tiles, n_tiles);
      const int t = tc.mt * kBlockM + row;
      const int n0 = tc.nt * p.block_n;
      ChunkStore st;
      st.alloc = p.out_alloc;
      st.n_store = p.n_store;
      st.row_in = t < p.M;
      const long long row_flat = (long long)t * p.out_ld + p.out_shift;
      const long long boff = (long long)tc.b * p.out_bstride;
      float* out0f = reinterpret_cast<float*>(p.out0) + boff;
      __nv_bfloat16* out0h = reinterpret_cast<__nv_bfloat16*>(p.out0) + boff;
      __nv_bfloat16* out1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + boff;

      if (tile_skipped(p, tc)) {
        if (p.zero_skipped) {  // keep "everything past the valid length is zero" true for final outputs
          st.valid = 0;
          float zero[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) zero[i] = 0.f;
          for (int c = g; c < n_chunks; c += 4) {
            const int n = n0 + c * 16;
            if (p.out0_dtype == OUT_F32) st.f32(out0f, row_flat + n, n, zero);
            else if (p.out0_dtype == OUT_BF16) st.bf16(out0h, row_flat + n, n, zero);
            if (p.out1_mode != OUT1_NONE) st.bf16(out1, row_flat + n, n, zero);
          }
        }
        continue;
      }
      st.valid = p.lengths ? min((long long)p.lengths[tc.b] * p.out_valid_mul, p.out_alloc) : p.out_alloc;
      const float* addf = reinterpret_cast<const float*>(p.addend) + boff;
      const __nv_bfloat16* addh = reinterpret_cast<const __nv_bfloat16*>(p.addend) + boff;
      const float* temb = p.temb ? p.temb + (long long)tc.b * p.temb_bstride : nullptr;
      float2* red = red_base + parity * (2 * 4 * kBlockM);
      parity ^= 1;

      // ---- accumulators -> registers in one burst, then hand the TMEM stage back to the MMA warp
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride);
      float v[4][16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (g + 4 * j < n_chunks) tmem_ld16(taddr + (uint32_t)((g + 4 * j) * 16), reinterpret_cast<uint32_t(&)[16]>(v[j]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;

      // ---- bias + pointwise activation, in place
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (g + 4 * j < n_chunks) {
          const int ch0 = (n0 + (g + 4 * j) * 16) % p.chan_mod;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + ch0) + g4);
            float* x = &v[j][4 * g4];
            x[0] += bv.x, x[1] += bv.y, x[2] += bv.z, x[3] += bv.w;
            if (p.act == ACT_LRELU || p.act == ACT_LRELU_TANH) {
#pragma unroll
              for (int i = 0; i < 4; ++i) x[i] = x[i] > 0.f ? x[i] : 0.1f * x[i];
              if (p.act == ACT_LRELU_TANH) {
#pragma unroll
                for (int i = 0; i < 4; ++i) x[i] = tanhf(x[i]);
              }
            } else if (p.act == ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 4; ++i) x[i] = gelu_erf(x[i]);
            }
          }
        }
      }

      if (p.act == ACT_LN_MISH) {  // LayerNorm over the 256-wide row, then Mish
        float mean, rstd;
        row_stats(v, red, g, row, mean, rstd);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(p.ln_g + (g + 4 * j) * 16) + g4);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(p.ln_b + (g + 4 * j) * 16) + g4);
            float* x = &v[j][4 * g4];
            x[0] = mish_f(fmaf((x[0] - mean) * rstd, gm.x, bt.x));
            x[1] = mish_f(fmaf((x[1] - mean) * rstd, gm.y, bt.y));
            x[2] = mish_f(fmaf((x[2] - mean) * rstd, gm.z, bt.z));
            x[3] = mish_f(fmaf((x[3] - mean) * rstd, gm.w, bt.w));
          }
        }
      }

      // ---- time-embedding add, residual add, primary store, copy / snake secondary store
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (g + 4 * j < n_chunks) {
          const int n = n0 + (g + 4 * j) * 16;
          const long long flat = row_flat + n;
          const bool partial = n + 16 > p.n_store;  // only the padded final conv (n_store = 1)
          const int stt = st.state(flat, partial ? max(p.n_store - n, 1) : 16);
          if (temb) {
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const float4 tv = __ldg(reinterpret_cast<const float4*>(temb + n) + g4);
              float* x = &v[j][4 * g4];
              x[0] += tv.x, x[1] += tv.y, x[2] += tv.z, x[3] += tv.w;
            }
          }
          if (p.addend && stt == 2) {
            if (p.addend_dtype == OUT_F32) {
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {
                const float4 a = *reinterpret_cast<const float4*>(addf + flat + 4 * g4);
                float* x = &v[j][4 * g4];
                x[0] += a.x, x[1] += a.y, x[2] += a.z, x[3] += a.w;
              }
            } else {
#pragma unroll
              for (int g8 = 0; g8 < 2; ++g8) {
                const uint4 a = *reinterpret_cast<const uint4*>(addh + flat + 8 * g8);
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                  v[j][8 * g8 + 2 * k] += __low2float(h2);
                  v[j][8 * g8 + 2 * k + 1] += __high2float(h2);
                }
              }
            }
          }
          if (stt != 0 && n < p.n_store) {
            if (partial) {  // scalar tail
              for (int i = 0; i < 16 && n + i < p.n_store; ++i) {
                const float x = stt == 2 ? v[j][i] : 0.f;
                if (p.out0_dtype == OUT_F32) out0f[flat + i] = x;
                else if (p.out0_dtype == OUT_BF16) out0h[flat + i] = __float2bfloat16(x);
                if (p.out1_mode == OUT1_COPY) out1[flat + i] = __float2bfloat16(x);
              }
            } else {
              if (p.out0_dtype == OUT_F32) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                  const float* x = &v[j][4 * g4];
                  reinterpret_cast<float4*>(out0f + flat)[g4] =
                      stt == 2 ? make_float4(x[0], x[1], x[2], x[3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              } else if (p.out0_dtype == OUT_BF16) {
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  const float* x = &v[j][8 * g8];
                  uint4 o = make_uint4(0u, 0u, 0u, 0u);
                  if (stt == 2)
                    o = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                   pack_bf16x2(x[6], x[7]));
                  reinterpret_cast<uint4*>(out0h + flat)[g8] = o;
                }
              }
              if (p.out1_mode == OUT1_COPY) {
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  const float* x = &v[j][8 * g8];
                  uint4 o = make_uint4(0u, 0u, 0u, 0u);
                  if (stt == 2)
                    o = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                   pack_bf16x2(x[6], x[7]));
                  reinterpret_cast<uint4*>(out1 + flat)[g8] = o;
                }
              } else if (p.out1_mode == OUT1_SNAKE) {
                const int ch0 = n % p.chan_mod;
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  const float* x = &v[j][8 * g8];
                  uint4 o = make_uint4(0u, 0u, 0u, 0u);
                  if (stt == 2) {
                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.p1_a + ch0) + 2 * g8);
                    const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.p1_a + ch0) + 2 * g8 + 1);
                    const float4 i0 = __ldg(reinterpret_cast<const float4*>(p.p1_b + ch0) + 2 * g8);
                    const float4 i1 = __ldg(reinterpret_cast<const float4*>(p.p1_b + ch0) + 2 * g8 + 1);
                    o.x = pack_bf16x2(snake_f(x[0], a0.x, i0.x), snake_f(x[1], a0.y, i0.y));
                    o.y = pack_bf16x2(snake_f(x[2], a0.z, i0.z), snake_f(x[3], a0.w, i0.w));
                    o.z = pack_bf16x2(snake_f(x[4], a1.x, i1.x), snake_f(x[5], a1.y, i1.y));
                    o.w = pack_bf16x2(snake_f(x[6], a1.z, i1.z), snake_f(x[7], a1.w, i1.w));
                  }
                  reinterpret_cast<uint4*>(out1 + flat)[g8] = o;
                }
              }
            }
          }
        }
      }

      if (p.out1_mode == OUT1_LN) {  // second output = LayerNorm(u), bf16 operand of the next GEMM
        float mean, rstd;
        row_stats(v, red + 4 * kBlockM, g, row, mean, rstd);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = n0 + (g + 4 * j) * 16;
          const long long flat = row_flat + n;
          const int stt = st.state(flat);
          if (stt == 0) continue;
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const float* x = &v[j][8 * g8];
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (stt == 2) {
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.p1_a + (g + 4 * j) * 16) + 2 * g8);
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.p1_a + (g + 4 * j) * 16) + 2 * g8 + 1);
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.p1_b + (g + 4 * j) * 16) + 2 * g8);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.p1_b + (g + 4 * j) * 16) + 2 * g8 + 1);
              o.x = pack_bf16x2(fmaf((x[0] - mean) * rstd, a0.x, b0.x), fmaf((x[1] - mean) * rstd, a0.y, b0.y));
              o.y = pack_bf16x2(fmaf((x[2] - mean) * rstd, a0.z, b0.z), fmaf((x[3] - mean) * rstd, a0.w, b0.w));
              o.z = pack_bf16x2(fmaf((x[4] - mean) * rstd, a1.x, b1.x), fmaf((x[5] - mean) * rstd, a1.y, b1.y));
              o.w = pack_bf16x2(fmaf((x[6] - mean) * rstd, a1.z, b1.z), fmaf((x[7] - mean) * rstd, a1.w, b1.w));
            }
            reinterpret_cast<uint4*>(out1 + flat)[g8] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + kRedBytes;
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
  const double kt = p.k_true > 0 ? p.k_true : p.kb_per_tap * kBlockK;
  const double rows = (double)p.B * p.M;
  const double out_b = (p.out0_dtype == OUT_F32 ? 4.0 : p.out0_dtype == OUT_BF16 ? 2.0 : 0.0) +
                       (p.out1_mode != OUT1_NONE ? 2.0 : 0.0) +
                       (p.addend ? (p.addend_dtype == OUT_F32 ? 4.0 : 2.0) : 0.0);
  ProfScope prof(stream, p.tag == 1 ? PK_CONV_DAC : PK_CONV_FLOW, 2.0 * rows * p.N * kt * p.taps,
                 rows * kt * 2.0 + (double)p.taps * p.N * kt * 2.0 + rows * (p.n_store < p.N ? p.n_store : p.N) * out_b);
  conv_gemm_kernel<<<grid, kThreads, smem, stream>>>(mapA0, mapA1, mapW, p, stages, tmem_cols, acc_stride);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ls
