#!/bin/bash
# first-block QKV as a conv_gemm launch over block2's LayerNorm output (LS_HEAD_VIA_CONV=1) vs the fused kernel's head launch (=0)
export LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_hv.so
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_boundary_gpu.py tests/test_ops_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -3
for d in 0 1 0 1; do
  LS_HEAD_VIA_CONV=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('head_via_conv=$d step ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
