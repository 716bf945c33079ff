// Micro-benchmark (development aid): per-SM TMA load throughput from L2 into shared memory.
// Every CTA streams `iters` boxes (64 bf16 x rows, 128B swizzle) of a [rows_total][K] bf16 matrix through a ring of
// `depth` slots with no consumer; prints bytes / clk / SM for shared (all CTAs read the same matrix) and private data.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../minimax-speech_b200/csrc/ptx.cuh"
using namespace ls;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) k_load(const __grid_constant__ CUtensorMap map, int box_rows, int depth,
                                                 int iters, int rows_total, int k_blocks, int private_rows,
                                                 long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slot_bytes = box_rows * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)depth * slot_bytes);
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int row_base = private_rows ? blockIdx.x * private_rows : 0;
    const int row_span = private_rows ? private_rows : rows_total;
    const int boxes_per_col = row_span / box_rows;
    const long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
      const int slot = it % depth;
      if (it >= depth) mbar_wait(&bars[slot], ((it / depth) - 1) & 1);
      if (it < iters) {
        const int kb = it % k_blocks, rb = (it / k_blocks) % boxes_per_col;
        mbar_arrive_expect_tx(&bars[slot], slot_bytes);
        tma_load_2d(smem + (size_t)slot * slot_bytes, &map, &bars[slot], kb * 64, row_base + rb * box_rows);
      }
    }
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}

int main(int argc, char** argv) {
  const int K = 1024;
  const int rows_total = 148 * 1024;  // 148 x 2 MB private regions; shared mode uses the first 1024 rows (2 MB)
  void* buf;
  cudaMalloc(&buf, (size_t)rows_total * K * 2);
  cudaMemset(buf, 1, (size_t)rows_total * K * 2);
  long long* out;
  cudaMalloc(&out, 148 * 8);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int iters = 2048;
  for (int box_rows : {64, 128, 256}) {
    for (int depth : {2, 5, 8, 12}) {
      if ((size_t)depth * box_rows * 128 > 200 * 1024) continue;
      for (int mode = 0; mode < 3; ++mode) {  // 0: all CTAs read the same 2 MB, 1: private 2 MB each, 2: one CTA only
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        cuuint32_t estr[2] = {1, 1};
        enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int grid = mode == 2 ? 1 : 148;
        const size_t smem = (size_t)depth * box_rows * 128 + 1024 + 256;
        for (int rep = 0; rep < 2; ++rep)
          k_load<<<grid, 128, smem>>>(map, box_rows, depth, iters, 1024, K / 64, mode == 1 ? 1024 : 0, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        double mx = 0, sum = 0;
        for (int i = 0; i < grid; ++i) { sum += h[i]; if (h[i] > mx) mx = h[i]; }
        const double bytes = (double)iters * box_rows * 128;
        printf("box_rows %3d depth %2d mode %s: %.1f B/clk/SM (avg), %.1f (slowest CTA); in flight %d KB\n", box_rows, depth,
               mode == 0 ? "shared " : mode == 1 ? "private" : "single ", bytes / (sum / grid), bytes / mx,
               depth * box_rows * 128 / 1024);
      }
    }
  }
  return 0;
}
