// Flash-attention forward for the estimator's self-attention (8 heads x 64), sm_100a.
//
// Replaces diffusers' Attention/AttnProcessor2_0 call inside BasicTransformerBlock
// (speech/matcha/models/components/transformer.py:196-204,266-271) together with the dense additive mask the
// reference materialises per block group (flow/decoder.py:441-445, utils/mask.py:161-236, utils/common.py:160-168):
// the key-padding mask is a per-utterance key bound, the streaming block-causal mask (chunk 50) a per-row bound.
//
// One CTA = 128 query rows of one (batch row, head); two CTAs share an SM (256 TMEM columns each).  Both GEMMs run
// with their A operand in TENSOR MEMORY (tcgen05.mma "TS" form): measured on B200 (profiles/micro/mma_bw.cu), an
// M = 128, N = 64 MMA whose A operand comes from shared memory occupies the tensor pipe for 72 clk (floor 32: the
// 4 KB A read is not hidden), 46 clk with A in TMEM -- with A in shared memory this kernel was bound by exactly that
// (3 CTAs x 8 MMAs x 72 clk per key tile against 1536 clk of MUFU work), not by the exponentials.
//   Q   TMA -> smem once, copied by its row-owning threads into TMEM (32 columns of packed bf16 pairs)
//   S   = Q K^T, fp32, double buffered in TMEM (S0 / S1): S(j+2) is issued right after P(j) V(j)
//   P   = exp2(S*c - m) as packed bf16 pairs, written by tcgen05.st INTO the first 32 columns of the S buffer it came
//         from (never through shared memory), and read from there as the A operand of O += P V
//   O   fp32 in TMEM, accumulated by the tensor core across key tiles, rescaled lazily (only when a row maximum grows
//         by more than 2^8) by the row's own thread
// Four softmax warps own one query row per thread (no cross-thread exchange).  The control warp issues TMA loads
// (K: 3 stages, V: 2 stages) and the MMAs; tcgen05.mma instructions execute in issue order, which is what makes the
// S(j+2)-over-P(j) reuse of a buffer safe.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "kernels.h"
#include "profiler.h"
#include "ptx.cuh"

namespace ls {
namespace {

constexpr int kQ = 128;
constexpr int kKV = ATTN_KV;       // keys per tile
static_assert(kKV == 64, "the TMEM layout below is laid out for 64-key tiles");
constexpr int kChunks = kKV / 32;  // 32-score register chunks per row and tile
constexpr int kCtasPerSm = 2;      // 256 TMEM columns per CTA
constexpr int kD = 64;
constexpr int kTile = kQ * kD * 2;  // 16 KB: 128 rows x 128 B
constexpr int kSoftmaxWarps = 4;    // warp w owns TMEM lanes [32w, 32w+32) = query rows
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
constexpr int kAttnThreads = kSoftmaxThreads + 32;  // + control warp
constexpr int kKVTile = kKV * kD * 2;  // one K or V tile: kKV rows x 128 B
constexpr int kKStages = 3, kVStages = 2;
constexpr int kAttnSmem = 2 * kTile + (kKStages + kVStages) * kKVTile + 1024 + 256;  // Q, K ring, V ring, output staging, align slack, barriers
constexpr int kAttnTmemCols = 256;
constexpr uint32_t kTmemS0 = 0, kTmemS1 = 64, kTmemO = 128, kTmemQ = 192;  // S0/S1: 64 fp32 columns (P: the first 32)
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays below 2^8 between rescales
#ifndef ATTN_POLY_MASK
#define ATTN_POLY_MASK 0x00
#endif
#ifndef ATTN_PACK_ALU
#define ATTN_PACK_ALU 0
#endif
constexpr int kPolyExpMask = ATTN_POLY_MASK;  // of every 8 score pairs, the ones whose exp2 runs on the FMA pipe

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
#if ATTN_PACK_ALU
    // bf16 pair without F2FP (an XU-pipe instruction, the pipe the exponentials already saturate): round half up on the
    // integer pipe (P is finite and non-negative), then one byte permute picks the two upper halves
    packed[i >> 1] = __byte_perm(__float_as_uint(a) + 0x8000u, __float_as_uint(b) + 0x8000u, 0x7632);
#else
    packed[i >> 1] = LS_PACK_H2(a, b);
#endif
  }
}

// One unit of work: 128 query rows of one (batch row, head).  Items are numbered with the query tile fastest, so CTAs that
// run at the same time read the same K / V out of L2.
struct Item {
  int b, h, q0, nkv, len;
};
__device__ __forceinline__ Item decode_item(const AttnParams& p, int it, int n_qt) {
  Item w;
  const int qt = it % n_qt;
  const int rest = it / n_qt;
  w.h = rest % p.H;
  w.b = rest / p.H;
  w.q0 = qt * kQ;
  int len = p.lengths ? p.lengths[w.b] : p.T;
  if (len > p.T) len = p.T;
  w.len = len;
  int tile_limit = len;
  if (p.chunk > 0) tile_limit = min(len, ((w.q0 + kQ - 1) / p.chunk + 1) * p.chunk);
  // a query tile that is padding only does no work: its rows are masked downstream
  w.nkv = w.q0 >= len ? 0 : (tile_limit + kKV - 1) / kKV;
  return w;
}
// Walks the key tiles of this CTA's items in order (items without work are skipped): the K and V loads run ahead of the
// MMAs along the same sequence, across item boundaries.
struct TileCursor {
  int it, j;
  Item w;
  __device__ __forceinline__ void seek(const AttnParams& p, int n_items, int n_qt, int stride) {  // first item with work at or after `it`
    j = 0;
    while (it < n_items) {
      w = decode_item(p, it, n_qt);
      if (w.nkv > 0) return;
      it += stride;
    }
  }
  __device__ __forceinline__ bool valid(int n_items) const { return it < n_items; }
  __device__ __forceinline__ void advance(const AttnParams& p, int n_items, int n_qt, int stride) {
    if (++j < w.nkv) return;
    it += stride;
    seek(p, n_items, n_qt, stride);
  }
};

template <bool kBias>
__global__ void __launch_bounds__(kAttnThreads, kCtasPerSm)
attn_kernel(const __grid_constant__ CUtensorMap mapQKV, const __grid_constant__ CUtensorMap mapOut,
            const __grid_constant__ AttnParams p) {
  pdl_launch_dependents();
  long long* tl = (p.timeline && blockIdx.x < 148) ? p.timeline + (size_t)blockIdx.x * 64 : nullptr;
#define TL(i)                    \
  do {                           \
    if (tl) tl[(i)] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) TL(0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;              // Q tile of the current item (free for the next item's load once copied to TMEM)
  uint8_t* sK = smem + kTile;      // kKStages stages
  uint8_t* sV = sK + kKStages * kKVTile;  // kVStages stages
  uint8_t* sOut = sV + kVStages * kKVTile;  // output staging tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + kTile);
  uint64_t* bar_q = bars;        // TMA -> softmax warps: Q tile in smem
  uint64_t* bar_qt = bars + 1;   // softmax warps -> MMA: Q copied to TMEM
  uint64_t* bar_k = bars + 2;    // [kKStages] TMA -> MMA
  uint64_t* bar_v = bars + 5;    // [kVStages] TMA -> MMA
  uint64_t* bar_s = bars + 7;    // [2] MMA -> softmax: S of a key tile in S buffer (tile & 1)
  uint64_t* bar_p = bars + 9;    // softmax -> MMA: P in TMEM, O rescaled
  uint64_t* bar_o = bars + 10;   // MMA -> softmax / loader: P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == kSoftmaxWarps) {
    if (lane == 0) {
      prefetch_tmap(&mapQKV);
      prefetch_tmap(&mapOut);
      mbar_init(bar_q, 1);
      mbar_init(bar_qt, kSoftmaxWarps);
      for (int i = 0; i < kKStages; ++i) mbar_init(&bar_k[i], 1);
      for (int i = 0; i < kVStages; ++i) mbar_init(&bar_v[i], 1);
      mbar_init(&bar_s[0], 1);
      mbar_init(&bar_s[1], 1);
      mbar_init(bar_p, kSoftmaxWarps);
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
  if (threadIdx.x == 0) TL(15);
  // Everything above (barriers, TMEM) overlapped the previous kernel's tail; lengths and QKV are read from here on.
  pdl_wait();
  // Persistent CTA: items blockIdx.x, blockIdx.x + gridDim.x, ...  Every barrier parity below is derived from running
  // counters (items started, key tiles consumed), identical in the control warp and the softmax warps.
  const int n_qt = (p.T + kQ - 1) / kQ;
  const int n_items = n_qt * p.H * p.B;
  const int stride = (int)gridDim.x;

  if (warp == kSoftmaxWarps) {
    // ------------------------------------------------ control warp: TMA loads + MMA issue.  The whole warp walks the
    // loop in uniform control flow (descriptors, coordinates and TMEM addresses stay in uniform registers); one elected
    // lane issues (see elect_one() in ptx.cuh).
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int inner = p.H * kD;
    TileCursor kc, vc, cur;  // K-load cursor, V-load cursor, item cursor
    kc.it = vc.it = cur.it = (int)blockIdx.x;
    kc.seek(p, n_items, n_qt, stride);
    vc = kc, cur = kc;
    int kl = 0, vl = 0;  // K / V tiles loaded so far (stage = count % stages)
    auto load_k = [&]() {  // next K tile of the sequence -> stage kl % kKStages
      if (!kc.valid(n_items)) return;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_k[kl % kKStages], kKVTile);
        tma_load_3d(sK + (kl % kKStages) * kKVTile, &mapQKV, &bar_k[kl % kKStages], inner + kc.w.h * kD, kc.j * kKV, kc.w.b);
      }
      __syncwarp();
      ++kl;
      kc.advance(p, n_items, n_qt, stride);
    };
    auto load_v = [&]() {
      if (!vc.valid(n_items)) return;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bar_v[vl & 1], kKVTile);
        tma_load_3d(sV + (vl & 1) * kKVTile, &mapQKV, &bar_v[vl & 1], 2 * inner + vc.w.h * kD, vc.j * kKV, vc.w.b);
      }
      __syncwarp();
      ++vl;
      vc.advance(p, n_items, n_qt, stride);
    };
    auto load_q = [&](const Item& w) {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, kTile);  // (the tensor map's boxes have kKV rows: the Q tile takes 128 / kKV loads)
#pragma unroll
        for (int i = 0; i < kQ / kKV; ++i) tma_load_3d(sQ + i * kKVTile, &mapQKV, bar_q, w.h * kD, w.q0 + i * kKV, w.b);
      }
      __syncwarp();
    };
    const uint32_t idesc_s = make_idesc_bf16(kQ, kKV, false, false);
    const uint32_t idesc_o = make_idesc_bf16(kQ, kD, false, true);  // B = V tile, MN-major
    int g = 0;           // key tiles whose P V has been issued (running over items)
    int sg = 0;          // key tiles whose S has been issued
    uint32_t n_item = 0; // items started
    // S of the next tile of the sequence into S buffer sg & 1: A = Q in TMEM (8 columns per 16-wide K step), B = K tile
    auto issue_s = [&]() {
      mbar_wait(&bar_k[sg % kKStages], (uint32_t)((sg / kKStages) & 1));
      tc_fence_after();
      const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + (sg % kKStages) * kKVTile));
      const uint32_t ds = tm + ((sg & 1) ? kTmemS1 : kTmemS0);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_bf16_ts(ds, tm + kTmemQ + 8 * k, dk + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&bar_s[sg & 1]);
      }
      __syncwarp();
      ++sg;
    };
    if (cur.valid(n_items)) {
      load_q(cur.w);
      load_k();
      load_v();
      load_k();
      load_k();
      load_v();
    }
    if (lane == 0) TL(1);
    while (cur.valid(n_items)) {
      const int nkv = cur.w.nkv;
      mbar_wait(bar_qt, n_item & 1);  // this item's Q sits in TMEM (and the previous item's O has been read out)
      tc_fence_after();
      ++n_item;
      if (lane == 0 && n_item == 1) TL(2);
      // the Q tile in shared memory is free: fetch the next item's
      TileCursor nxt = cur;
      nxt.it += stride;
      nxt.seek(p, n_items, n_qt, stride);
      if (nxt.valid(n_items)) load_q(nxt.w);
      issue_s();
      if (nkv > 1) issue_s();
      for (int j = 0; j < nkv; ++j, ++g) {
        mbar_wait(bar_p, g & 1);  // P written into S buffer g & 1, O rescaled where needed
        tc_fence_after();
        if (g < 4 && lane == 0) TL(3 + 2 * g);
        mbar_wait(&bar_v[g & 1], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        {
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + (g & 1) * kKVTile));
          const uint32_t dp = tm + ((g & 1) ? kTmemS1 : kTmemS0);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kKV / 16; ++kk)  // V rows of 16 keys are 2048 B apart (>> 4 = 128 in the descriptor)
              umma_bf16_ts(tm + kTmemO, dp + 8 * kk, dv + 128 * kk, idesc_o, (j | kk) != 0 ? 1u : 0u);
            umma_commit(bar_o);
          }
          __syncwarp();
        }
        if (g < 4 && lane == 0) TL(4 + 2 * g);
        // S(j+2) overwrites the buffer P(j) sits in: it is issued after P(j) V(j), and the tensor pipe runs in order
        if (j + 2 < nkv) issue_s();
        // the K stage of this tile is free (its S retired before the softmax warps could read it): the K ring runs
        // kKStages tiles ahead, into the next items
        load_k();
        mbar_wait(bar_o, g & 1);  // P V of this tile retired: its V stage is free
        load_v();
      }
      cur = nxt;
    }
  } else {
    // ------------------------------------------------ softmax warps: one thread per query row
    const int r = warp * 32 + lane;
    const float c = p.scale_log2e;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint32_t o_addr = tmem_base + kTmemO + lane_addr;
    const int sw = r & 7;
    long long* tls = threadIdx.x == 0 ? tl : nullptr;
#define TLS(i)                     \
  do {                             \
    if (tls) tls[(i)] = clock64(); \
  } while (0)
    TileCursor cur;
    cur.it = (int)blockIdx.x;
    cur.seek(p, n_items, n_qt, stride);
    int g = 0;            // key tiles consumed (running over items): S buffer and barrier parities
    uint32_t n_item = 0;  // items started
    bool stored = false;  // this warp's lane 0 has an output store in flight
    while (cur.valid(n_items)) {
      const Item w = cur.w;
      const int nkv = w.nkv;
      const int qi = w.q0 + r;
      int limit = w.len;
      if (p.chunk > 0) limit = min(w.len, (qi / p.chunk + 1) * p.chunk);
      {  // this row of Q: 128 B of the swizzled smem tile -> 32 TMEM columns (packed bf16 pairs, K order)
        mbar_wait(bar_q, n_item & 1);
        ++n_item;
        uint32_t qv[32];
        const uint8_t* qrow = sQ + r * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 v = *reinterpret_cast<const uint4*>(qrow + ((ch ^ sw) << 4));
          qv[4 * ch + 0] = v.x, qv[4 * ch + 1] = v.y, qv[4 * ch + 2] = v.z, qv[4 * ch + 3] = v.w;
        }
        tmem_st32(tmem_base + kTmemQ + lane_addr, qv);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_qt);
      }
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < nkv; ++j, ++g) {
        const uint32_t s_addr = tmem_base + ((g & 1) ? kTmemS1 : kTmemS0) + lane_addr;
        mbar_wait(&bar_s[g & 1], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (g < 4) TLS(16 + 6 * g);
        uint32_t s[kChunks][32];
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc) tmem_ld32(s_addr + cc * 32, s[cc]);
        tmem_ld_wait();
        if (g < 4) TLS(17 + 6 * g);
        if (kBias) {  // relative-position term: this row's 64 consecutive entries of its (skewed) bias row
          const int qc = qi < p.T ? qi : p.T - 1;
          const float* br = p.bias + ((long long)w.b * p.H + w.h) * p.bias_bh + (long long)qc * p.bias_ld + (p.T - 1 - qc) + j * kKV;
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
        if (g < 4) TLS(18 + 6 * g);
        const float mt = mx * c;
        const bool grow = mt > m + kRescaleThreshold;  // false for mt = -inf; true for m = -inf and finite mt
        const float m_new = grow ? mt : m;
        const float alpha = (grow && j > 0) ? ex2_approx(m - m_new) : 1.0f;
        const float m_use = m_new == -INFINITY ? 0.f : m_new;
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pk[kChunks * 16];
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc)
          exp_chunk(s[cc], c, m_use, reinterpret_cast<uint32_t(&)[16]>(pk[cc * 16]), sum0, sum1);
        l = fmaf(l, alpha, sum0 + sum1);
        m = m_new;
        if (g < 4) TLS(19 + 6 * g);

        if (j > 0) {
          mbar_wait(bar_o, (g - 1) & 1);  // P V of the previous tile retired: O is stable
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
          }
        }
        if (g < 4) TLS(20 + 6 * g);
        // P: 64 keys = 32 packed columns, over the first half of the S buffer the scores came from
        tmem_st32(s_addr, reinterpret_cast<const uint32_t(&)[32]>(pk));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p);
        if (g < 4) TLS(21 + 6 * g);
      }

      mbar_wait(bar_o, (g - 1) & 1);
      tc_fence_after();
      if (n_item == 1) TLS(40);
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      // O / l -> bf16 -> this row's 128-byte line of the staging tile, 128B-swizzled; each warp then hands its 32 rows to
      // one TMA store (rows past T are clipped by the tensor map).  A direct store would make every warp instruction
      // touch 32 different rows with 16 B each.
      if (stored) {  // the previous item's store has read this warp's rows of the staging tile
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
      }
      uint8_t* orow = sOut + r * 128;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld32(o_addr + hh * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          uint4 v;
          v.x = LS_PACK_H2(__uint_as_float(o[8 * gq + 0]) * inv, __uint_as_float(o[8 * gq + 1]) * inv);
          v.y = LS_PACK_H2(__uint_as_float(o[8 * gq + 2]) * inv, __uint_as_float(o[8 * gq + 3]) * inv);
          v.z = LS_PACK_H2(__uint_as_float(o[8 * gq + 4]) * inv, __uint_as_float(o[8 * gq + 5]) * inv);
          v.w = LS_PACK_H2(__uint_as_float(o[8 * gq + 6]) * inv, __uint_as_float(o[8 * gq + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (((hh * 4 + gq) ^ sw) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (w.q0 + warp * 32 < p.T) {
        if (lane == 0) {
          tma_store_3d(&mapOut, sOut + warp * 32 * 128, w.h * kD, w.q0 + warp * 32, w.b);
          bulk_commit();
        }
        stored = true;
      }
      if (n_item == 1) TLS(41);
      cur.it += stride;
      cur.seek(p, n_items, n_qt, stride);
    }
    if (stored && lane == 0) bulk_wait_read<0>();  // the CTA may exit once the source has been read
  }

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
  const long long n_items = (long long)((p.T + kQ - 1) / kQ) * p.H * p.B;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // LS_ATTN_PERSISTENT=1: two persistent CTAs per SM walk the items (no per-item launch gap / TMEM allocation / barrier
  // setup, K / V prefetched across items); default: one CTA per item, scheduled by the hardware as slots free up --
  // with 3.5 items per slot at the bench shape the hardware's dynamic balance beats the persistent form's 4-item slots
  static const bool persistent = [] {
    const char* e = getenv("LS_ATTN_PERSISTENT");
    return e && e[0] == '1';
  }();
  dim3 grid((unsigned)(persistent ? std::min<long long>(n_items, (long long)kCtasPerSm * sms) : n_items));
  const double bh = (double)p.B * p.H;
  ProfScope prof(stream, PK_ATTENTION, 4.0 * bh * p.T * (double)p.T * kD, bh * p.T * kD * 2.0 * 4.0);
  count_launch();
  AttnParams pp = p;
  pp.timeline = (g_debug_buffer && g_debug_bytes >= 148 * 64 * 8) ? g_debug_buffer : nullptr;
  pp.timeline_entries = g_debug_bytes / 8;
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
