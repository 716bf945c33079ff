// Micro-benchmark 4 (development aid): cost of one TMA load / one tcgen05.mma issued from `lane == 0` code versus from
// an elect.sync-guarded block with warp-uniform operands.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../minimax-speech_b200/csrc/ptx.cuh"
using namespace ls;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <bool kElect>
__global__ void __launch_bounds__(128, 1) k_load(const __grid_constant__ CUtensorMap map, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kDepth = 4, kSlot = 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDepth * kSlot);
  if (threadIdx.x == 0) {
    prefetch_tmap(&map);
    for (int i = 0; i < kDepth; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    if (kElect) {
      int slot = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {  // whole warp walks the loop, one elected lane issues
        if (it >= kDepth) mbar_wait(&bars[slot], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[slot], kSlot);
          tma_load_3d(smem + slot * kSlot, &map, &bars[slot], 0, (it & 7) * 128, it & 15);
        }
        __syncwarp();
        if (++slot == kDepth) slot = 0, ph ^= 1;
      }
    } else if (threadIdx.x == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        if (it >= kDepth) mbar_wait(&bars[slot], ph ^ 1);
        mbar_arrive_expect_tx(&bars[slot], kSlot);
        tma_load_3d(smem + slot * kSlot, &map, &bars[slot], 0, (it & 7) * 128, it & 15);
        if (++slot == kDepth) slot = 0, ph ^= 1;
      }
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
}

// MMA issue cost: 64 x (4 MMAs of 128x128x16 + commit) on garbage smem data, accumulators in TMEM
template <bool kElect>
__global__ void __launch_bounds__(128, 1) k_mma(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t idesc = make_idesc_bf16(128, 128, false, false);
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    if (kElect) {
      for (int it = 0; it < iters; ++it) {
        const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + (it & 1) * 16384));
        const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 32768 + (it & 1) * 16384));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem + (it & 1) * 128, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
          umma_commit(bar);
        }
        __syncwarp();
      }
    } else if (threadIdx.x == 0) {
      for (int it = 0; it < iters; ++it) {
        const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + (it & 1) * 16384));
        const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + 32768 + (it & 1) * 16384));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + (it & 1) * 128, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
        umma_commit(bar);
      }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) {
      // wait for the last commit (phase parity of arrival #iters)
      mbar_wait(bar, (iters - 1) & 1);
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

int main() {
  const int K = 1024, rows = 1024;
  uint8_t* buf;
  cudaMalloc(&buf, (size_t)rows * K * 2);
  cudaMemset(buf, 0, (size_t)rows * K * 2);
  long long* out;
  cudaMalloc(&out, 148 * 2 * 8);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(k_load<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_load<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 1024;
  long long h[4];
  for (int e = 0; e < 2; ++e) {
    for (int rep = 0; rep < 2; ++rep) {
      if (e) k_load<true><<<1, 128, 70 * 1024>>>(map, iters, out);
      else k_load<false><<<1, 128, 70 * 1024>>>(map, iters, out);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("load error\n"); return 1; }
    cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("TMA load 16 KB, %s: %.1f clk per load\n", e ? "elect.sync, warp-uniform" : "lane == 0 branch       ", (double)h[0] / iters);
  }
  for (int e = 0; e < 2; ++e) {
    for (int rep = 0; rep < 2; ++rep) {
      if (e) k_mma<true><<<1, 128, 70 * 1024>>>(iters, out);
      else k_mma<false><<<1, 128, 70 * 1024>>>(iters, out);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("mma error\n"); return 1; }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("4 x tcgen05.mma 128x128x16 + commit, %s: issue %.1f clk per group, complete %.1f clk per group (tensor time 256)\n",
           e ? "elect.sync, warp-uniform" : "lane == 0 branch       ", (double)h[0] / iters, (double)h[1] / iters);
  }
  return 0;
}
