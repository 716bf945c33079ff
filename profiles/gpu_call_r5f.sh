#!/bin/bash
# L2 prefetch of a launch's weight boxes in the prologue of tblock / estimator conv_gemm (default) vs none (nopf.so)
mkdir -p gpurun_out
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
NOPF="-DTBLOCK_L2_PREFETCH=0 -DCONV_L2_PREFETCH=0"
echo "=== kernel + parity tests (default)"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_front_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -2
for n in default nopf default nopf; do
  case $n in default) d="";; nopf) d="$NOPF";; esac
  run $n "$d" timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
