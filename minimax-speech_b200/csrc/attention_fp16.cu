// fp16-operand build of the flash-attention kernel (see the LS_HALF_FP16 note in ptx.cuh).
#define LS_HALF_FP16 1
#include "attention.cu"
