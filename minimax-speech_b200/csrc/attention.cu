// Flash-attention forward for the estimator's self-attention (8 heads x 64), sm_100a.
//
// Replaces diffusers' Attention/AttnProcessor2_0 call inside BasicTransformerBlock
// (speech/matcha/models/components/transformer.py:196-204,266-271) together with the dense additive mask the
// reference materialises per block group (flow/decoder.py:441-445, utils/mask.py:161-236, utils/common.py:160-168):
// the key-padding mask is a per-utterance key bound, the streaming block-causal mask (chunk 50) a per-row bound.
//
// One CTA = 128 query rows of one (batch row, head).  S = Q K^T and O_j = P V run on tcgen05 with fp32
// accumulators in TMEM; 8 softmax warps (two threads per query row, 64 keys each, row max / sum exchanged through
// smem) run the online softmax, write P as bf16 into a 128B-swizzled K-major smem tile and keep the running
// output in registers.  Two CTAs share an SM so one CTA's MMAs overlap the other's softmax.
#include <cmath>

#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

constexpr int kQ = 128;
constexpr int kKV = 128;
constexpr int kD = 64;
constexpr int kTile = kQ * kD * 2;  // 16 KB: 128 rows x 128 B
constexpr int kSoftmaxWarps = 8;                   // warp w: TMEM lane quarter w%4, key half w/4
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
constexpr int kAttnThreads = kSoftmaxThreads + 32;  // + control warp
constexpr int kAttnSmem = 6 * kTile + 1024 + 128 + 3 * 1024;  // Q, K x2, V, P x2, align slack, barriers, exchange
constexpr int kAttnTmemCols = 256;                 // S: cols [0,128)   O_j: cols [128,192)

__device__ __forceinline__ void softmax_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxThreads) : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ AttnParams p) {
  const int q0 = blockIdx.x * kQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  int len = p.lengths ? p.lengths[b] : p.T;
  if (len > p.T) len = p.T;
  if (q0 >= len) return;  // query tile is padding only: its rows are masked downstream
  int tile_limit = len;
  if (p.chunk > 0) tile_limit = min(len, ((q0 + kQ - 1) / p.chunk + 1) * p.chunk);
  const int nkv = (tile_limit + kKV - 1) / kKV;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;      // 2 stages
  uint8_t* sV = smem + 3 * kTile;
  uint8_t* sP = smem + 4 * kTile;  // 2 K-blocks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTile);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  float* s_max = reinterpret_cast<float*>(bars + 16);  // [2 parity][2 half][128]
  float* s_sum = s_max + 2 * 2 * kQ;                   // [2 half][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == kSoftmaxWarps) {
    if (lane == 0) {
      prefetch_tmap(&mapQKV);
      mbar_init(bar_q, 1);
      mbar_init(&bar_k[0], 1);
      mbar_init(&bar_k[1], 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, kSoftmaxThreads);
      mbar_init(bar_o, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kAttnTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == kSoftmaxWarps) {
    if (lane == 0) {
      // ------------------------------------------------ control thread: TMA loads + MMA issue
      const int inner = p.H * kD;
      const int colq = h * kD, colk = inner + h * kD, colv = 2 * inner + h * kD;
      mbar_arrive_expect_tx(bar_q, kTile);
      tma_load_3d(sQ, &mapQKV, bar_q, colq, q0, b);
      mbar_arrive_expect_tx(&bar_k[0], kTile);
      tma_load_3d(sK, &mapQKV, &bar_k[0], colk, 0, b);
      mbar_arrive_expect_tx(bar_v, kTile);
      tma_load_3d(sV, &mapQKV, bar_v, colv, 0, b);
      if (nkv > 1) {
        mbar_arrive_expect_tx(&bar_k[1], kTile);
        tma_load_3d(sK + kTile, &mapQKV, &bar_k[1], colk, kKV, b);
      }
      const uint32_t idesc_s = make_idesc_bf16(kQ, kKV, false, false);
      const uint32_t idesc_o = make_idesc_bf16(kQ, kD, false, true);  // B = V tile, MN-major
      const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + (j & 1) * kTile));
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(bar_s);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(bar_p, j & 1);
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kKV / 16; ++kk) {
          const uint64_t dp = make_smem_desc_sw128(smem_u32(sP + (kk >> 2) * kTile)) + 2 * (kk & 3);
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + kk * 2048));
          umma_bf16(tmem_o, dp, dv, idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(bar_o);
        if (j + 1 < nkv) {
          mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
          mbar_wait(bar_o, j & 1);  // P V_j retired: V, P and K_j buffers are free
          mbar_arrive_expect_tx(bar_v, kTile);
          tma_load_3d(sV, &mapQKV, bar_v, colv, (j + 1) * kKV, b);
          if (j + 2 < nkv) {
            mbar_arrive_expect_tx(&bar_k[j & 1], kTile);
            tma_load_3d(sK + (j & 1) * kTile, &mapQKV, &bar_k[j & 1], colk, (j + 2) * kKV, b);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax warps: 2 threads per query row (64 keys each)
    const int quarter = warp & 3;
    const int half = warp >> 2;
    const int r = quarter * 32 + lane;
    const int qi = q0 + r;
    int limit = len;
    if (p.chunk > 0) limit = min(len, (qi / p.chunk + 1) * p.chunk);
    const float c = p.scale_log2e;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_s + lane_addr + (uint32_t)(half * 64);
    const uint32_t o_addr = tmem_o + lane_addr + (uint32_t)(half * 32);
    float m = -INFINITY, l = 0.f;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    uint8_t* prow = sP + half * kTile + r * 128;  // this thread's 64 keys = one 128-byte swizzled row
    const int sw = r & 7;

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      const int nvalid = limit - j * kKV - half * 64;  // valid keys among this thread's 64
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t s[32];
        tmem_ld32(s_addr + cc * 32, s);
        tmem_ld_wait();
        if (nvalid >= 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(s[i]));
        }
      }
      float* xm = s_max + (j & 1) * 2 * kQ;
      xm[half * kQ + r] = mx;
      softmax_barrier();
      mx = fmaxf(mx, xm[(half ^ 1) * kQ + r]);
      const float m_new = fmaxf(m, mx * c);
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ex2_approx(m - m_use);
      float rowsum = 0.f;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t s[32];
        tmem_ld32(s_addr + cc * 32, s);
        tmem_ld_wait();
        float pv[32];
        if (nvalid >= 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = ex2_approx(fmaf(__uint_as_float(s[i]), c, -m_use));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = ex2_approx(fmaf(__uint_as_float(s[i]), c, -m_use));
            pv[i] = cc * 32 + i < nvalid ? e : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) rowsum += pv[i];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 v;
          v.x = pack_bf16x2(pv[8 * q4 + 0], pv[8 * q4 + 1]);
          v.y = pack_bf16x2(pv[8 * q4 + 2], pv[8 * q4 + 3]);
          v.z = pack_bf16x2(pv[8 * q4 + 4], pv[8 * q4 + 5]);
          v.w = pack_bf16x2(pv[8 * q4 + 6], pv[8 * q4 + 7]);
          const int ch = cc * 4 + q4;
          *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = v;
        }
      }
      l = fmaf(l, alpha, rowsum);
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_p);

      mbar_wait(bar_o, j & 1);
      tc_fence_after();
      {
        uint32_t s[32];
        tmem_ld32(o_addr, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(s[i]));
      }
    }
    s_sum[half * kQ + r] = l;
    softmax_barrier();
    l += s_sum[(half ^ 1) * kQ + r];
    if (qi < p.T) {
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      __nv_bfloat16* dst = p.out + ((long long)b * p.T + qi) * (p.H * kD) + h * kD + half * 32;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * g + 0] * inv, o[8 * g + 1] * inv);
        v.y = pack_bf16x2(o[8 * g + 2] * inv, o[8 * g + 3] * inv);
        v.z = pack_bf16x2(o[8 * g + 4] * inv, o[8 * g + 5] * inv);
        v.w = pack_bf16x2(o[8 * g + 6] * inv, o[8 * g + 7] * inv);
        *reinterpret_cast<uint4*>(dst + 8 * g) = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
}

}  // namespace

cudaError_t launch_attention(const CUtensorMap& mapQKV, const AttnParams& p, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  dim3 grid((p.T + kQ - 1) / kQ, p.H, p.B);
  const double bh = (double)p.B * p.H;
  ProfScope prof(stream, PK_ATTENTION, 4.0 * bh * p.T * (double)p.T * kD, bh * p.T * kD * 2.0 * 4.0);
  attn_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(mapQKV, p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ls
