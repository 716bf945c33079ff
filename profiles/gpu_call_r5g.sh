#!/bin/bash
# next-launch weight prefetch hints (conv weights asked for by the launch before them; default) vs the previous library (prev.so)
run() { local name=$1; shift
  if [ "$name" = default ]; then "$@"; else LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/prev.so "$@"; fi; }
echo "=== parity tests (default)"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_boundary_gpu.py tests/test_ops_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -2
for n in default prev default prev; do
  run $n timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
