// Micro-benchmark 5 (development aid, round 2): is the fused transformer-block kernel's weight stream (every CTA pulls the
// SAME 16 KB boxes of a 2 MB weight set out of L2 in the same order at about the same time) bound by the handful of L2
// slices that hold the box everybody wants, rather than by bandwidth or by the ring depth?
//
// grid CTAs (one per SM) stream `passes` x 2 MB through a ring of `depth` slots with `nprod` producer warps (the tblock
// structure: load i is issued by warp i % nprod, slot i % depth) and one consumer warp that hands every slot back after
// `delay` clocks (0 = pure streaming; 256 = the time the MMAs of a 16 KB box take).  Modes:
//   0  every CTA reads the same matrix in the same order               (what tblock does today)
//   1  same matrix, every CTA starts at a different box (rotation)     (changes the order of K blocks: numerics!)
//   2  R replicas of the matrix in memory, CTA c reads replica c % R   (same order, same numerics, R x the L2 footprint)
//   3  every CTA has its own 512 KB (private data: the no-sharing bound)
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
  int depth, nprod, box_rows, delay, mode, replicas, passes, kcols, rows;
};

// matrix: [replica][rows][kcols] bf16, box = 64 columns x box_rows rows; 3-D map (64-col K block index, row, replica*...)
__global__ void __launch_bounds__(256, 1) k_stream(const __grid_constant__ CUtensorMap map, const P p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int slot_bytes = p.box_rows * 128;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.depth * slot_bytes);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    prefetch_tmap(&map);
    for (int i = 0; i < p.depth; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kbs = p.kcols / 64, row_boxes = p.rows / p.box_rows;
  const int n_boxes = kbs * row_boxes;             // boxes per pass
  const int total = n_boxes * p.passes;
  int rot = 0, rep = 0;
  if (p.mode == 1) rot = (int)(((long long)blockIdx.x * n_boxes) / gridDim.x);
  if (p.mode == 2) rep = blockIdx.x % p.replicas;
  if (p.mode == 3) rep = blockIdx.x;
  const long long t0 = clock64();
  if (warp < p.nprod) {
    // incremental slot / phase / box bookkeeping (no divisions on the issue path)
    int slot = warp % p.depth, use = warp / p.depth;
    int bi = (warp + rot) % n_boxes;
    const int step_slot = p.nprod % p.depth, step_use = p.nprod / p.depth;
    for (int i = warp; i < total; i += p.nprod) {
      mbar_wait(&empty[slot], (use & 1) ^ 1);
      if (elect_one()) {
        const int kb = bi % kbs, rb = bi / kbs;
        mbar_arrive_expect_tx(&full[slot], slot_bytes);
        tma_load_3d(smem + (size_t)slot * slot_bytes, &map, &full[slot], kb * 64, rb * p.box_rows, rep);
      }
      __syncwarp();
      slot += step_slot, use += step_use;
      if (slot >= p.depth) slot -= p.depth, use += 1;
      bi += p.nprod;
      if (bi >= n_boxes) bi -= n_boxes;
    }
  } else if (warp == p.nprod) {
    int slot = 0;
    uint32_t ph = 0;
    for (int i = 0; i < total; ++i) {
      mbar_wait(&full[slot], ph);
      if (p.delay > 0) {
        const long long t = clock64();
        while (clock64() - t < p.delay) {
        }
      }
      if (lane == 0) mbar_arrive(&empty[slot]);
      __syncwarp();
      if (++slot == p.depth) slot = 0, ph ^= 1;
    }
    if (lane == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const size_t cap = (size_t)160 << 20;  // 160 MB arena (private mode: 148 x 512 KB = 74 MB)
  void* buf;
  cudaMalloc(&buf, cap);
  cudaMemset(buf, 1, cap);
  long long* out;
  cudaMalloc(&out, 148 * 8);
  struct Cfg {
    int kcols, rows, box_rows, depth, nprod, delay, mode, replicas, grid;
  };
  std::vector<Cfg> cfgs;
  for (int grid : {125, 148})
    for (int kcols : {256, 1024}) {
      const int rows = 2 * 1024 * 1024 / (kcols * 2);  // 2 MB
      for (int delay : {0, 256}) {
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 0, 1, grid});
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 1, 1, grid});
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 2, 2, grid});
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 2, 4, grid});
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 2, 8, grid});
        cfgs.push_back({kcols, rows, 128, 5, 3, delay, 2, 16, grid});
        cfgs.push_back({kcols, rows / 4, 128, 5, 3, delay, 3, 1, grid});
      }
      // ring shape at delay 0: deeper ring, bigger boxes, more producers
      cfgs.push_back({kcols, rows, 128, 10, 3, 0, 0, 1, grid});
      cfgs.push_back({kcols, rows, 128, 10, 5, 0, 0, 1, grid});
      cfgs.push_back({kcols, rows, 256, 5, 3, 0, 0, 1, grid});
      cfgs.push_back({kcols, rows, 256, 5, 3, 0, 2, 4, grid});
      cfgs.push_back({kcols, rows, 128, 10, 5, 0, 2, 4, grid});
      cfgs.push_back({kcols, rows, 64, 10, 5, 0, 0, 1, grid});
    }
  cfgs.push_back({256, 4096, 128, 5, 3, 0, 0, 1, 1});
  cfgs.push_back({256, 4096, 128, 5, 3, 0, 0, 1, 16});
  cfgs.push_back({256, 4096, 128, 5, 3, 0, 0, 1, 64});
  for (const Cfg& c : cfgs) {
    const int reps = c.mode == 3 ? 148 : c.replicas;
    const size_t mat_bytes = (size_t)c.rows * c.kcols * 2;
    if (mat_bytes * reps > cap) { printf("skip (arena)\n"); continue; }
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)c.kcols, (cuuint64_t)c.rows, (cuuint64_t)reps};
    cuuint64_t strides[2] = {(cuuint64_t)c.kcols * 2, (cuuint64_t)mat_bytes};
    cuuint32_t box[3] = {64, (cuuint32_t)c.box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    P p{c.depth, c.nprod, c.box_rows, c.delay, c.mode, c.replicas, c.mode == 3 ? 16 : 4, c.kcols, c.rows};
    const size_t smem = (size_t)c.depth * c.box_rows * 128 + 1024 + 1024 + (c.depth * c.box_rows * 128 < 120 * 1024 ? 120 * 1024 : 0);
    double best = 0, best_slow = 0;
    for (int it = 0; it < 3; ++it) {
      k_stream<<<c.grid, 256, smem>>>(map, p, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, out, sizeof(long long) * c.grid, cudaMemcpyDeviceToHost);
      double mx = 0, sum = 0;
      for (int i = 0; i < c.grid; ++i) { sum += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
      const double bytes = (double)mat_bytes * p.passes;
      const double avg = bytes / (sum / c.grid), slow = bytes / mx;
      if (avg > best) best = avg, best_slow = slow;
    }
    printf("K %4d box %3d rows depth %2d prod %d delay %3d mode %d rep %2d grid %3d : %6.1f B/clk/SM avg, %6.1f slowest CTA\n",
           c.kcols, c.box_rows, c.depth, c.nprod, c.delay, c.mode, c.replicas, c.grid, best, best_slow);
    fflush(stdout);
  }
  return 0;
}
