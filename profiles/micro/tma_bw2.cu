// Micro-benchmark 2 (development aid): what does one TMA instruction cost?  3-D boxes (64 x rows x kblocks), several
// issuing warps, issue-only timing.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../minimax-speech_b200/csrc/ptx.cuh"
using namespace ls;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// each of `nwarps` warps (lane 0) runs its own ring of `depth` slots
__global__ void __launch_bounds__(256, 1) k_load(const __grid_constant__ CUtensorMap map, int slot_bytes, int depth,
                                                 int iters, int box_rows, int box_kb, int nwarps, int issue_only,
                                                 long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nwarps * depth * slot_bytes);
  if (threadIdx.x == 0) {
    prefetch_tmap(&map);
    for (int i = 0; i < nwarps * depth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if (w < nwarps && (threadIdx.x & 31) == 0) {
    uint8_t* my = smem + (size_t)w * depth * slot_bytes;
    uint64_t* mb = bars + w * depth;
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int slot = it % depth;
      if (!issue_only && it >= depth) mbar_wait(&mb[slot], ((it / depth) - 1) & 1);
      if (it < iters) {
        const int kb = (it * box_kb) % 16, rb = (it / (16 / box_kb) + w * 3) % (1024 / box_rows);
        mbar_arrive_expect_tx(&mb[slot], slot_bytes);
        tma_load_3d(my + (size_t)slot * slot_bytes, &map, &mb[slot], 0, rb * box_rows, kb);
      }
    }
    const long long t1 = clock64();
    if (issue_only)
      for (int s = 0; s < depth; ++s) mbar_wait(&mb[s], 0);
    out[blockIdx.x * 8 + w] = t1 - t0;
  }
}

int main() {
  const int K = 1024, rows = 1024;  // 2 MB matrix, L2 resident, read by every CTA
  void* buf;
  cudaMalloc(&buf, (size_t)rows * K * 2);
  cudaMemset(buf, 1, (size_t)rows * K * 2);
  long long* out;
  cudaMalloc(&out, 148 * 8 * 8);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int box_rows, box_kb, depth, nwarps, issue_only, iters; };
  const Cfg cfgs[] = {
      {128, 1, 4, 1, 0, 1024}, {128, 1, 4, 2, 0, 1024}, {128, 1, 2, 4, 0, 1024}, {128, 1, 1, 8, 0, 1024},
      {128, 2, 4, 1, 0, 1024}, {128, 4, 2, 1, 0, 1024}, {128, 4, 3, 1, 0, 1024}, {256, 2, 3, 1, 0, 1024},
      {256, 4, 2, 1, 0, 512},  {64, 1, 8, 1, 0, 1024},  {128, 1, 8, 1, 1, 8},    {128, 4, 3, 1, 1, 3},
      {128, 1, 8, 4, 1, 8},    {32, 1, 8, 1, 0, 1024},  {8, 1, 8, 1, 0, 1024},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap map;
    // dims: 64 elements of one K block | rows | K blocks
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
    cuuint32_t box[3] = {64, (cuuint32_t)c.box_rows, (cuuint32_t)c.box_kb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const int slot_bytes = c.box_rows * 128 * c.box_kb;
    const size_t smem = (size_t)c.nwarps * c.depth * slot_bytes + 1024 + 1024;
    if (smem > 227 * 1024) { printf("skip (smem)\n"); continue; }
    for (int grid : {1, 148}) {
      for (int rep = 0; rep < 2; ++rep)
        k_load<<<grid, 256, smem>>>(map, slot_bytes, c.depth, c.iters, c.box_rows, c.box_kb, c.nwarps, c.issue_only, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148 * 8];
      cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i)
        for (int w = 0; w < c.nwarps; ++w) if (h[i * 8 + w] > mx) mx = h[i * 8 + w];
      const double bytes = (double)c.iters * slot_bytes * c.nwarps;
      printf("box %3d rows x %d kb (%3d KB) depth %d warps %d %s grid %3d: %6.1f B/clk/SM, %7.1f clk per instruction%s\n",
             c.box_rows, c.box_kb, slot_bytes / 1024, c.depth, c.nwarps, c.issue_only ? "issue-only" : "ring      ", grid,
             bytes / mx, mx / c.iters, c.issue_only ? " (issue cost)" : "");
    }
  }
  return 0;
}
