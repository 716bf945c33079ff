// Micro-benchmark 3 (development aid): TMA issue cost per lane vs per warp, and 1-D bulk copies (no tensor map).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../minimax-speech_b200/csrc/ptx.cuh"
using namespace ls;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mode 0: `nlanes` lanes of warp 0 each run a ring (tensor loads); mode 1: lane 0 of `nlanes` warps, 1-D bulk copies
__global__ void __launch_bounds__(256, 1) k_load(const __grid_constant__ CUtensorMap map, const uint8_t* src, int slot_bytes,
                                                 int depth, int iters, int nlanes, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nlanes * depth * slot_bytes);
  if (threadIdx.x == 0) {
    prefetch_tmap(&map);
    for (int i = 0; i < nlanes * depth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int id = mode == 0 ? lane : w;
  const bool active = mode == 0 ? (w == 0 && lane < nlanes) : (lane == 0 && w < nlanes);
  if (active) {
    uint8_t* my = smem + (size_t)id * depth * slot_bytes;
    uint64_t* mb = bars + id * depth;
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int slot = it % depth;
      if (it >= depth) mbar_wait(&mb[slot], ((it / depth) - 1) & 1);
      if (it < iters) {
        mbar_arrive_expect_tx(&mb[slot], slot_bytes);
        if (mode == 0)
          tma_load_3d(my + (size_t)slot * slot_bytes, &map, &mb[slot], 0, ((it + id * 3) % 8) * 128, it % 16);
        else
          bulk_load_1d(my + (size_t)slot * slot_bytes, src + (size_t)((it + id * 5) % 24) * slot_bytes, slot_bytes, &mb[slot]);
      }
    }
    out[blockIdx.x * 32 + id] = clock64() - t0;
  }
}

int main() {
  const int K = 1024, rows = 1024;
  uint8_t* buf;
  cudaMalloc(&buf, (size_t)rows * K * 2);
  cudaMemset(buf, 1, (size_t)rows * K * 2);
  long long* out;
  cudaMalloc(&out, 148 * 32 * 8);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  CUtensorMap map;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  struct Cfg { int mode, nlanes, depth, slot_bytes, iters; };
  const Cfg cfgs[] = {{0, 1, 4, 16384, 1024}, {0, 2, 4, 16384, 1024}, {0, 4, 2, 16384, 1024}, {0, 8, 1, 16384, 1024},
                      {1, 1, 4, 16384, 1024}, {1, 1, 2, 65536, 1024}, {1, 1, 8, 4096, 1024},  {1, 4, 2, 16384, 1024},
                      {1, 2, 1, 65536, 1024}, {1, 1, 12, 16384, 1024}};
  for (const Cfg& c : cfgs) {
    const size_t smem = (size_t)c.nlanes * c.depth * c.slot_bytes + 2048;
    if (smem > 227 * 1024) { printf("skip (smem)\n"); continue; }
    for (int grid : {1, 148}) {
      for (int rep = 0; rep < 2; ++rep) k_load<<<grid, 256, smem>>>(map, buf, c.slot_bytes, c.depth, c.iters, c.nlanes, c.mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      static long long h[148 * 32];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i)
        for (int w = 0; w < c.nlanes; ++w) if (h[i * 32 + w] > mx) mx = h[i * 32 + w];
      const double bytes = (double)c.iters * c.slot_bytes * c.nlanes;
      printf("%s x%d, %2d KB per instruction, depth %2d, grid %3d: %6.1f B/clk/SM, %7.1f clk per instruction per issuer\n",
             c.mode == 0 ? "tensor load, lanes of one warp" : "1-D bulk copy, warps          ", c.nlanes, c.slot_bytes / 1024, c.depth,
             grid, bytes / mx, mx / c.iters);
    }
  }
  return 0;
}
