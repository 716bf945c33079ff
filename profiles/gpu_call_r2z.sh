#!/bin/bash
# res_conv fused into the resnet conv2 launch (second accumulator): parity, then same-library A/B (LS_CONV_FUSE_RES=0 / 1)
export LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_fr.so
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -4
for d in 0 1 0 1; do
  LS_CONV_FUSE_RES=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fuse=$d step ms', round(d['ms_per_step'],2), 'launches', d['gpu_launches'], {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
