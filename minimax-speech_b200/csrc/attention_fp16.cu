// fp16-operand build of the flash-attention kernel (see the LS_HALF_FP16 note in ptx.cuh).
#ifndef LS_NO_FP16_BUILD
#define LS_HALF_FP16 1
#include "attention.cu"
#else  // development aid: a library without the fp16-operand kernels
#include "kernels.h"
namespace ls {
cudaError_t launch_attention_fp16(const CUtensorMap&, const AttnParams&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace ls
#endif
