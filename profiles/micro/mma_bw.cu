// Micro-benchmark 6 (development aid, round 2): what does one tcgen05.mma (M = 128, bf16) cost, alone and while the
// shared memory is busy with TMA fills and LDS / STS traffic?  SS mode (A and B from shared memory) reads
// (128 + N) x 32 B per K = 16 step; TS mode (A from tensor memory) only N x 32 B.
//   warp 0        issues `iters` MMAs back to back into one accumulator, commits, waits; clk / MMA reported
//   warps 1..3    (bg & 1) each streams 16 KB boxes of a 2 MB matrix into its own two slots as fast as they land
//   warps 4..11   (bg & 2) conflict-free STS.128 + LDS.128 loop over a private 4 KB region per warp
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../minimax-speech_b200/csrc/ptx.cuh"
using namespace ls;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct P {
  int mode;   // 0 SS, 1 TS
  int n;      // 128 or 256
  int iters, bg;
  int pattern;  // 0: MMAs only; 1: tcgen05.commit after every 4 MMAs (conv_gemm frees a weight slot per tap); 2: + a wait on an
                // already-complete mbarrier and the fence before them; 3: pattern 2 with every MMA group under its own election
  int shift;  // A descriptor starts `shift` rows into the 128B-swizzled tile (conv_gemm's halo taps), base offset = shift & 7
};

constexpr int kTile = 16384;
// smem: A tile (16 KB) | B tiles 2 x 32 KB | TMA slots 3 warps x 2 x 16 KB | LSU regions 8 x 4 KB | barriers
constexpr int kOffB = kTile, kOffTma = kOffB + 2 * 32768, kOffLsu = kOffTma + 6 * kTile, kOffBar = kOffLsu + 8 * 4096;
constexpr int kSmem = kOffBar + 256 + 1024;

__global__ void __launch_bounds__(384, 1) k_mma(const __grid_constant__ CUtensorMap map, const P p, long long* out,
                                                volatile int* dummy) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* done = bars;          // MMA completion
  uint64_t* tfull = bars + 1;     // [6] TMA slots
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kOffBar / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    stop = 0;
    prefetch_tmap(&map);
    mbar_init(done, 1);
    for (int i = 0; i < 6; ++i) mbar_init(&tfull[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, p.n, false, false);
    const uint64_t a0 = make_smem_desc_sw128(smem_u32(smem) + (uint32_t)p.shift * 128u, (uint32_t)p.shift & 7u);
    const uint64_t b0 = make_smem_desc_sw128(smem_u32(smem + kOffB));
    const uint64_t b1 = make_smem_desc_sw128(smem_u32(smem + kOffB + 32768));
    const uint32_t d = tmem, at = tmem + 256;
    __syncwarp();
    const long long t0 = clock64();
    if (p.pattern) {  // conv_gemm's per-tap shape: [wait ready barrier] 4 MMAs, commit to a slot barrier
      uint64_t* slot = tfull;       // commit target (never waited on)
      uint64_t* ready = tfull + 1;  // completed once below, waited on with its completed parity
      if (lane == 0) mbar_arrive(ready);
      __syncwarp();
      if (p.pattern < 3) {
        if (elect_one()) {
          for (int it = 0; it < p.iters; it += 4) {
            if (p.pattern >= 2) {
              mbar_wait(ready, 0);
              tc_fence_after();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            umma_commit(slot);
          }
        }
        __syncwarp();
      } else {
        for (int it = 0; it < p.iters; it += 4) {
          mbar_wait(ready, 0);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, 1u);
            umma_commit(slot);
          }
          __syncwarp();
        }
      }
    } else
    for (int it = 0; it < p.iters; it += 8) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t bd = ((k & 4) ? b1 : b0) + 2 * (k & 3);
          if (p.mode == 0) umma_bf16(d, a0 + 2 * (k & 3), bd, idesc, 1u);
          else umma_bf16_ts(d, at + 8 * (k & 7), bd, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(done, 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 2] = t2 - t0;
      out[blockIdx.x * 2 + 1] = t1 - t0;
      *reinterpret_cast<volatile int*>(&stop) = 1;
    }
  } else if (warp <= 3) {
    if (p.bg & 1) {
      uint8_t* my = smem + kOffTma + (warp - 1) * 2 * kTile;
      uint64_t* mb = tfull + (warp - 1) * 2;
      int it = 0;
      uint32_t ph = 0;
      if (elect_one()) {
        for (int s = 0; s < 2; ++s) {
          mbar_arrive_expect_tx(&mb[s], kTile);
          tma_load_2d(my + s * kTile, &map, &mb[s], 0, ((warp * 7 + s) & 31) * 128);
        }
      }
      __syncwarp();
      long long bytes = 0;
      while (*reinterpret_cast<volatile int*>(&stop) == 0) {
        const int s = it & 1;
        mbar_wait(&mb[s], ph);
        if (elect_one()) {
          mbar_arrive_expect_tx(&mb[s], kTile);
          tma_load_2d(my + s * kTile, &map, &mb[s], (it & 3) * 64, ((warp * 7 + it) & 31) * 128);
        }
        __syncwarp();
        bytes += kTile;
        if (s) ph ^= 1;
        ++it;
      }
      mbar_wait(&mb[it & 1], ph);  // drain the two loads still in flight
      mbar_wait(&mb[(it + 1) & 1], (it & 1) ? ph ^ 1 : ph);
      if (lane == 0 && blockIdx.x == 0) out[300 + warp] = bytes;
    }
  } else {
    if (p.bg & 2) {
      uint4* r = reinterpret_cast<uint4*>(smem + kOffLsu + (warp - 4) * 4096);
      uint4 v = make_uint4(lane, warp, 0, 0);
      long long n = 0;
      while (*reinterpret_cast<volatile int*>(&stop) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          r[k * 32 + lane] = v;
          const uint4 w = r[((k + 3) & 7) * 32 + lane];
          v.x += w.y;
        }
        n += 8 * 2 * 512;
      }
      if (v.x == 0x7fffffff) *dummy = 1;
      if (lane == 0 && blockIdx.x == 0) out[310 + warp] = n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  void* buf;
  cudaMalloc(&buf, 2 << 20);
  cudaMemset(buf, 0, 2 << 20);
  long long* out;
  cudaMalloc(&out, 512 * 8);
  int* dummy;
  cudaMalloc(&dummy, 4);
  CUtensorMap map;
  cuuint64_t dims[2] = {256, 4096};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    printf("encode failed\n");
    return 1;
  }
  for (int grid : {1, 148})
    for (int mode = 0; mode < 2; ++mode)
      for (int n : {64, 128, 256})
        for (int bg = 0; bg < 4; ++bg) {
          P p{mode, n, 2048, bg, 0, 0};
          cudaMemset(out, 0, 512 * 8);
          for (int rep = 0; rep < 2; ++rep) k_mma<<<grid, 384, kSmem>>>(map, p, out, dummy);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s (mode %d n %d bg %d)\n", cudaGetErrorString(e), mode, n, bg); return 1; }
          long long h[512];
          cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
          double tot = 0, iss = 0;
          for (int i = 0; i < grid; ++i) tot += (double)h[2 * i], iss += (double)h[2 * i + 1];
          tot /= grid, iss /= grid;
          const double tma_b = (double)(h[301] + h[302] + h[303]), lsu_b = 0;
          double lsu = 0;
          for (int w = 4; w < 12; ++w) lsu += (double)h[310 + w];
          printf("grid %3d %s N=%3d bg=%d (tma %d, lsu %d): %6.1f clk/MMA (issue %5.1f); floor %3d; smem operand %5.1f B/clk; "
                 "bg tma %5.1f B/clk, bg lsu %5.1f B/clk\n",
                 grid, mode ? "TS" : "SS", n, bg, bg & 1, (bg >> 1) & 1, tot / p.iters, iss / p.iters, n / 2,
                 (mode ? n * 32.0 : (128 + n) * 32.0) / (tot / p.iters), tma_b / tot, lsu / tot);
          (void)lsu_b;
          fflush(stdout);
        }
  for (int n : {48, 64, 96, 128, 192, 256})
    for (int pattern = 0; pattern < 4; ++pattern) {
      P p{0, n, 2048, 0, pattern, 0};
      cudaMemset(out, 0, 512 * 8);
      for (int rep = 0; rep < 2; ++rep) k_mma<<<148, 384, kSmem>>>(map, p, out, dummy);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s (n %d pattern %d)\n", cudaGetErrorString(e), n, pattern); return 1; }
      long long h[512];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double tot = 0;
      for (int i = 0; i < 148; ++i) tot += (double)h[2 * i];
      printf("grid 148 SS N=%3d pattern %d (0 plain, 1 commit / 4 MMAs, 2 + ready wait, 3 + election per group): %6.1f clk/MMA; floor %3d\n", n, pattern,
             tot / 148 / p.iters, n / 2);
    }
  // row-shifted A descriptors (the halo taps of conv_gemm) at the DAC decoder's thin widths
  for (int n : {48, 96, 192})
    for (int shift : {0, 1, 4, 8, 9, 27}) {
      P p{0, n, 2048, 0, 0, shift};
      cudaMemset(out, 0, 512 * 8);
      for (int rep = 0; rep < 2; ++rep) k_mma<<<148, 384, kSmem>>>(map, p, out, dummy);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s (n %d shift %d)\n", cudaGetErrorString(e), n, shift); return 1; }
      long long h[512];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double tot = 0;
      for (int i = 0; i < 148; ++i) tot += (double)h[2 * i];
      printf("grid 148 SS N=%3d A shifted by %2d rows: %6.1f clk/MMA; floor %3d\n", n, shift, tot / 148 / p.iters, n / 2);
    }
  return 0;
}
