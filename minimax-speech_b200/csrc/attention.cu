// Flash-attention forward for the estimator's self-attention (8 heads x 64), sm_100a.
//
// Replaces diffusers' Attention/AttnProcessor2_0 call inside BasicTransformerBlock
// (speech/matcha/models/components/transformer.py:196-204,266-271) together with the dense additive mask the
// reference materialises per block group (flow/decoder.py:441-445, utils/mask.py:161-236, utils/common.py:160-168):
// the key-padding mask is a per-utterance key bound, the streaming block-causal mask (chunk 50) a per-row bound.
//
// One CTA = 128 query rows of one (batch row, head); two CTAs share an SM.  S = Q K^T and O += P V run on tcgen05 with
// fp32 accumulators in TMEM (S: columns [0,128), O: [128,192)).  Four softmax warps own one query row per thread
// (no cross-thread exchange): the 128 scores of a key tile are pulled into registers once, the row maximum only
// moves when it grows by more than 2^8 (so O, which stays in TMEM and is accumulated by the tensor core across
// key tiles, is rescaled rarely and only by rows that need it), P goes to a 128B-swizzled K-major smem tile.
// The control warp's single thread issues TMA loads and MMAs; S(j+1) is issued before P(j) V(j), so the next
// tile's scores are ready while the softmax warps are still writing P(j).
//
// Round 2 measured a second form of this kernel, kept as profiles/attention_ts_variant.cu.txt: both GEMMs in the tcgen05
// "TS" form (Q copied into tensor memory, P written by tcgen05.st over its S buffer and never through shared memory, S
// double buffered, 256 TMEM columns and two CTAs per SM), optionally with persistent CTAs (K / V prefetched across
// items) or with two threads per query row.  An M = 128, N = 64 MMA occupies the tensor pipe for 46 clk in TS form
// against 72 clk here (profiles/micro/mma_bw.cu), and alone that kernel is faster (35.3 against 37.7 us at B = 32,
// T = 500) -- but inside the bench step, back to back with the fused block kernel, it is slower (25.3 against 23.8 ms per
// step, same box): the softmax warps are bound by their own dependency chains (ncu: "wait" 2.3 and scoreboard 1.8 stall
// cycles per issued instruction, XU throttle 0.09; polynomial exp2 and integer-pipe bf16 packing change nothing), and
// three resident CTAs hide them better than two.  This three-CTA form therefore stays the product kernel.
#include <cmath>

#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

constexpr int kQ = 128;
constexpr int kKV = ATTN_KV;       // keys per tile: 128 (2 CTAs per SM) or 64 (3 CTAs per SM, 128 TMEM columns each)
constexpr int kChunks = kKV / 32;  // 32-score register chunks per row and tile
constexpr int kPBlocks = kKV / 64; // 64-key K blocks of the P tile
constexpr int kCtasPerSm = kKV == 128 ? 2 : 3;
constexpr int kD = 64;
constexpr int kTile = kQ * kD * 2;  // 16 KB: 128 rows x 128 B
constexpr int kSoftmaxWarps = 4;    // warp w owns TMEM lanes [32w, 32w+32) = query rows
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
constexpr int kAttnThreads = kSoftmaxThreads + 32;  // + control warp
constexpr int kKVTile = kKV * kD * 2;  // one K or V tile: kKV rows x 128 B
constexpr int kAttnSmem = kTile + 3 * kKVTile + kPBlocks * kTile + 1024 + 256;  // Q, K x2, V, P, align slack, barriers
constexpr int kAttnTmemCols = kKV == 128 ? 256 : 128;  // S: cols [0,kKV)   O: cols [kKV,kKV+64)
constexpr float kRescaleThreshold = 8.0f;
#ifndef ATTN_POLY_MASK
#define ATTN_POLY_MASK 0x00
#endif
constexpr int kPolyExpMask = ATTN_POLY_MASK;  // of every 8 score pairs, the ones whose exp2 runs on the FMA pipe           // log2 units: P stays below 2^8 between rescales

// p[i] = 2^(s[i]*c - m) for 32 scores (masked scores are -inf -> 0); returns the packed bf16 pairs and adds to the row sum
__device__ __forceinline__ void exp_chunk(const uint32_t (&s)[32], float c, float m, uint32_t (&packed)[16],
                                          float& sum0, float& sum1) {
  const float nm = -m;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float a, b;
    ffma2(a, b, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), c, c, nm, nm);
    if (kPolyExpMask & (1 << ((i >> 1) & 7))) {  // this pair on the FMA pipe
      ex2_poly2(a, b);
    } else {
      a = ex2_approx(a);
      b = ex2_approx(b);
    }
    fadd2(sum0, sum1, sum0, sum1, a, b);
    packed[i >> 1] = LS_PACK_H2(a, b);
  }
}

template <bool kBias>
__global__ void __launch_bounds__(kAttnThreads, kCtasPerSm)
attn_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapOut,
            const __grid_constant__ AttnParams p) {
  pdl_launch_dependents();
  const int q0 = blockIdx.x * kQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  long long* tl = (p.timeline && cta_lin < 148) ? p.timeline + (size_t)cta_lin * 64 : nullptr;
#define TL(i)                    \
  do {                           \
    if (tl) tl[(i)] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) TL(0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;      // 2 stages
  uint8_t* sV = sK + 2 * kKVTile;
  uint8_t* sP = sV + kKVTile;      // kPBlocks K-blocks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kPBlocks * kTile);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;  // MMA -> softmax: S(j) in TMEM
  uint64_t* bar_p = bars + 5;  // softmax -> MMA: P(j) in smem, S(j) consumed, O rescaled
  uint64_t* bar_o = bars + 6;  // MMA -> softmax / loader: P(j) V(j) retired
  uint64_t* bar_f = bars + 7;  // softmax -> MMA: S(j) is in registers, the S columns may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == kSoftmaxWarps) {
    if (lane == 0) {
      prefetch_tmap(&mapQKV);
      prefetch_tmap(&mapOut);
      mbar_init(bar_q, 1);
      mbar_init(&bar_k[0], 1);
      mbar_init(&bar_k[1], 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, kSoftmaxWarps);
      mbar_init(bar_o, 1);
      mbar_init(bar_f, kSoftmaxWarps);
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
  if (threadIdx.x == 0) TL(15);
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + kKV;
  // Everything above (barriers, TMEM) overlapped the previous kernel's tail; lengths and QKV are read from here on.
  pdl_wait();
  int len = p.lengths ? p.lengths[b] : p.T;
  if (len > p.T) len = p.T;
  int tile_limit = len;
  if (p.chunk > 0) tile_limit = min(len, ((q0 + kQ - 1) / p.chunk + 1) * p.chunk);
  // a query tile that is padding only does no work: its rows are masked downstream
  const int nkv = q0 >= len ? 0 : (tile_limit + kKV - 1) / kKV;

  if (nkv == 0) {
    // fall through to the common exit (TMEM is released there)
  } else if (warp == kSoftmaxWarps) {
    {
      // ------------------------------------------------ control warp: TMA loads + MMA issue.  The whole warp walks the
      // loop in uniform control flow (descriptors, coordinates and TMEM addresses stay in uniform registers); one elected
      // lane issues (see elect_one() in ptx.cuh).
      const uint32_t tmem_su = __shfl_sync(0xffffffffu, tmem_s, 0);
      const uint32_t tmem_ou = __shfl_sync(0xffffffffu, tmem_o, 0);
      const int nkv_u = __shfl_sync(0xffffffffu, nkv, 0);
      const int inner = p.H * kD;
      const int colq = h * kD, colk = inner + h * kD, colv = 2 * inner + h * kD;
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, kTile);  // (the tensor map's boxes have kKV rows: the Q tile takes 128 / kKV loads)
#pragma unroll
        for (int i = 0; i < kQ / kKV; ++i) tma_load_3d(sQ + i * kKVTile, &mapQKV, bar_q, colq, q0 + i * kKV, b);
        mbar_arrive_expect_tx(&bar_k[0], kKVTile);
        tma_load_3d(sK, &mapQKV, &bar_k[0], colk, 0, b);
        mbar_arrive_expect_tx(bar_v, kKVTile);
        tma_load_3d(sV, &mapQKV, bar_v, colv, 0, b);
        if (nkv_u > 1) {
          mbar_arrive_expect_tx(&bar_k[1], kKVTile);
          tma_load_3d(sK + kKVTile, &mapQKV, &bar_k[1], colk, kKV, b);
        }
      }
      __syncwarp();
      const uint32_t idesc_s = make_idesc_bf16(kQ, kKV, false, false);
      const uint32_t idesc_o = make_idesc_bf16(kQ, kD, false, true);  // B = V tile, MN-major
      const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + (j & 1) * kKVTile));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) umma_bf16(tmem_su, dq + 2 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(bar_s);
        }
        __syncwarp();
      };
      if (lane == 0) TL(1);
      mbar_wait(bar_q, 0);
      mbar_wait(&bar_k[0], 0);
      tc_fence_after();
      if (lane == 0) TL(2);
      issue_s(0);
      const uint64_t dp0 = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t dp1 = make_smem_desc_sw128(smem_u32(sP + kTile));
      const uint64_t dv0 = make_smem_desc_sw128(smem_u32(sV));
      for (int j = 0; j < nkv_u; ++j) {
        // the softmax warps hold S(j) in registers: the next scores can be computed while they work on this tile
        mbar_wait(bar_f, j & 1);
        tc_fence_after();
        if (j + 1 < nkv_u) {
          mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        if (j + 2 < nkv_u) {  // S(j) has retired (the softmax warps read it): K buffer j&1 is free
          if (elect_one()) {
            mbar_arrive_expect_tx(&bar_k[j & 1], kKVTile);
            tma_load_3d(sK + (j & 1) * kKVTile, &mapQKV, &bar_k[j & 1], colk, (j + 2) * kKV, b);
          }
          __syncwarp();
        }
        mbar_wait(bar_p, j & 1);  // P(j) written, O rescaled where needed
        tc_fence_after();
        if (j < 4 && lane == 0) TL(3 + 2 * j);
        mbar_wait(bar_v, j & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kKV / 16; ++kk)  // V rows of 16 keys are 2048 B apart (>> 4 = 128 in the descriptor)
            umma_bf16(tmem_ou, (kk < 4 ? dp0 : dp1) + 2 * (kk & 3), dv0 + 128 * kk, idesc_o, (j | kk) != 0 ? 1u : 0u);
          umma_commit(bar_o);
        }
        __syncwarp();
        if (j < 4 && lane == 0) TL(4 + 2 * j);
        if (j + 1 < nkv_u) {
          mbar_wait(bar_o, j & 1);  // P(j) V(j) retired: the V buffer is free
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_v, kKVTile);
            tma_load_3d(sV, &mapQKV, bar_v, colv, (j + 1) * kKV, b);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------ softmax warps: one thread per query row
    const int r = warp * 32 + lane;
    const int qi = q0 + r;
    int limit = len;
    if (p.chunk > 0) limit = min(len, (qi / p.chunk + 1) * p.chunk);
    const float c = p.scale_log2e;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint32_t s_addr = tmem_s + lane_addr;
    const uint32_t o_addr = tmem_o + lane_addr;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;  // 64 keys = one 128-byte swizzled row per K block
    const int sw = r & 7;

    long long* tls = threadIdx.x == 0 ? tl : nullptr;
#define TLS(i)                     \
  do {                             \
    if (tls) tls[(i)] = clock64(); \
  } while (0)
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, j & 1);
      tc_fence_after();
      if (j < 4) TLS(16 + 6 * j);
      uint32_t s[kChunks][32];
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) tmem_ld32(s_addr + cc * 32, s[cc]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_f);
      if (j < 4) TLS(17 + 6 * j);
      if (kBias) {  // relative-position term: this row's 64 consecutive entries of its (skewed) bias row
        const int qc = qi < p.T ? qi : p.T - 1;
        const float* br = p.bias + ((long long)b * p.H + h) * p.bias_bh + (long long)qc * p.bias_ld + (p.T - 1 - qc) + j * kKV;
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int jj = cc * 32 + i;
            const float bv = (j * kKV + jj < p.T) ? __ldg(br + jj) : 0.f;
            s[cc][i] = __float_as_uint(__uint_as_float(s[cc][i]) + bv);
          }
      }
      const int nvalid = limit - j * kKV;  // valid keys of this row in this tile (may be <= 0 for streaming rows)
      if (nvalid < kKV) {  // masked keys: -inf scores (exp2 -> 0); only the last key tile of a row pays for this
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc * 32 + i >= nvalid) s[cc][i] = 0xff800000u;
      }
      float pm[8];  // independent chains: the reduction is latency-, not issue-bound
#pragma unroll
      for (int a = 0; a < 8; ++a) pm[a] = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc)
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          pm[(i >> 1) & 7] = fmaxf(pm[(i >> 1) & 7], fmaxf(__uint_as_float(s[cc][i]), __uint_as_float(s[cc][i + 1])));
      const float mx = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));
      // lazy running maximum: move it only when it grows by more than the threshold (first tile: always)
      if (j < 4) TLS(18 + 6 * j);
      const float mt = mx * c;
      const bool grow = mt > m + kRescaleThreshold;  // false for mt = -inf; true for m = -inf and finite mt
      const float m_new = grow ? mt : m;
      const float alpha = (grow && j > 0) ? ex2_approx(m - m_new) : 1.0f;
      const float m_use = m_new == -INFINITY ? 0.f : m_new;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[kChunks][16];
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) exp_chunk(s[cc], c, m_use, pk[cc], sum0, sum1);
      l = fmaf(l, alpha, sum0 + sum1);
      m = m_new;
      if (j < 4) TLS(19 + 6 * j);

      if (j > 0) {
        mbar_wait(bar_o, (j - 1) & 1);  // P(j-1) V(j-1) retired: O is stable and the P tile may be overwritten
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld32(o_addr + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr + hh * 32, o);
          }
          tmem_st_wait();
        }
      }
      if (j < 4) TLS(20 + 6 * j);
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) {
        uint8_t* blk = prow + (cc >> 1) * kTile;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint4 v = make_uint4(pk[cc][4 * q4], pk[cc][4 * q4 + 1], pk[cc][4 * q4 + 2], pk[cc][4 * q4 + 3]);
          const int ch = (cc & 1) * 4 + q4;
          *reinterpret_cast<uint4*>(blk + ((ch ^ sw) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
      if (j < 4) TLS(21 + 6 * j);
    }

    mbar_wait(bar_o, (nkv - 1) & 1);
    tc_fence_after();
    TLS(40);
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    // O / l -> bf16 -> this row's 128-byte line of the (now idle) Q tile, 128B-swizzled; each warp then hands its 32
    // rows to one TMA store (rows past T are clipped by the tensor map).  A direct store would make every warp
    // instruction touch 32 different rows with 16 B each.
    uint8_t* orow = sQ + r * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o[32];
      tmem_ld32(o_addr + hh * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = LS_PACK_H2(__uint_as_float(o[8 * g + 0]) * inv, __uint_as_float(o[8 * g + 1]) * inv);
        v.y = LS_PACK_H2(__uint_as_float(o[8 * g + 2]) * inv, __uint_as_float(o[8 * g + 3]) * inv);
        v.z = LS_PACK_H2(__uint_as_float(o[8 * g + 4]) * inv, __uint_as_float(o[8 * g + 5]) * inv);
        v.w = LS_PACK_H2(__uint_as_float(o[8 * g + 6]) * inv, __uint_as_float(o[8 * g + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + (((hh * 4 + g) ^ sw) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && q0 + warp * 32 < p.T) {
      tma_store_3d(&mapOut, sQ + warp * 32 * 128, h * kD, q0 + warp * 32, b);
      bulk_commit();
      bulk_wait_read<0>();  // the CTA may exit (and its shared memory be reused) once the source has been read
    }
  }

  if (threadIdx.x == 0) TL(41);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TL(42);
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttnTmemCols);
  }
#undef TL
#undef TLS
}

}  // namespace

cudaError_t LS_FN(launch_attention)(const CUtensorMap& mapQKV, const AttnParams& p, cudaStream_t stream) {
#if !LS_HALF_FP16
  if (p.fp16) return launch_attention_fp16(mapQKV, p, stream);  // fp16-operand build of this file
#endif
  static std::atomic<unsigned long long> optin0{0}, optin1{0};  // one bit per device
  if (cudaError_t e = smem_optin_once(optin0, reinterpret_cast<const void*>(attn_kernel<false>), kAttnSmem); e != cudaSuccess) return e;
  if (cudaError_t e = smem_optin_once(optin1, reinterpret_cast<const void*>(attn_kernel<true>), kAttnSmem); e != cudaSuccess) return e;
  if (p.B <= 0 || p.T <= 0) return cudaSuccess;
  dim3 grid((p.T + kQ - 1) / kQ, p.H, p.B);
  const double bh = (double)p.B * p.H;
  ProfScope prof(stream, PK_ATTENTION, 4.0 * bh * p.T * (double)p.T * kD, bh * p.T * kD * 2.0 * 4.0);
  count_launch();
  AttnParams pp = p;
  pp.timeline = (g_debug_buffer && g_debug_bytes >= 148 * 64 * 8) ? g_debug_buffer : nullptr;
  // output tensor map ([B][T][H*64] bf16, 32-row boxes: one TMA store per softmax warp), cached per (buffer, shape)
  struct OutMap {
    const void* out = nullptr;
    int B = 0, T = 0, H = 0;
    CUtensorMap map;
  };
  static thread_local OutMap cache[8];
  static thread_local int next = 0;
  const OutMap* hit = nullptr;
  for (const OutMap& e : cache)
    if (e.out == p.out && e.B == p.B && e.T == p.T && e.H == p.H) hit = &e;
  if (!hit) {
    OutMap& e = cache[next];
    next = (next + 1) % 8;
    e.out = nullptr;
    if (!make_act_map(&e.map, p.out, p.H * kD, p.T, p.B, p.H * kD, (long long)p.T * p.H * kD, 32)) return cudaErrorInvalidValue;
    e.out = p.out, e.B = p.B, e.T = p.T, e.H = p.H;
    hit = &e;
  }
  if (p.bias) return launch_pdl(attn_kernel<true>, grid, dim3(kAttnThreads), (size_t)kAttnSmem, stream, 1, mapQKV, hit->map, pp);
  return launch_pdl(attn_kernel<false>, grid, dim3(kAttnThreads), (size_t)kAttnSmem, stream, 1, mapQKV, hit->map, pp);
}

}  // namespace ls
