// Optional per-launch timing (CUDA events on the launching stream) used by bench.py for the roofline figures.
#pragma once
#include <cuda_runtime.h>

namespace ls {

enum ProfKind : int { PK_CONV_FLOW = 0, PK_ATTENTION = 1, PK_CONV_DAC = 2, PK_ELEMENTWISE = 3, PK_TBLOCK = 4, PK_COUNT = 5 };

bool prof_enabled();
void prof_begin();
// aggregates since prof_begin(); synchronises the device.  arrays of PK_COUNT entries.
void prof_end(long long* launches, double* ms, double* flops, double* bytes);

class ProfScope {
 public:
  ProfScope(cudaStream_t s, int kind, double flops, double bytes);
  ~ProfScope();

 private:
  cudaStream_t s_;
  int slot_;
};

}  // namespace ls
