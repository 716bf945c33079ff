#!/bin/bash
# tblock: slot-release commits inside the MMA election (me1) vs separate elections (me0); same box, interleaved
export LS_NO_REBUILD=1
LS_LIB=$PWD/build_variants/libls_me1.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k tblock 2>&1 | tail -2
for pass in 1 2 3; do
  for v in me0 me1; do
    LS_LIB=$PWD/build_variants/libls_$v.so timeout 100 python -c "
import profiles.time_kernels as tk
print('$v', end=' '); tk.tblock(16000); print('$v', end=' '); tk.tblock(16000)"
  done
done
for v in me0 me1 me0 me1; do
  LS_LIB=$PWD/build_variants/libls_$v.so timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
