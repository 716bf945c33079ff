#!/bin/bash
# dual-issue mode restricted to streamed-weight launches: same-library A/B (LS_CONV_DUAL=0 / 1)
export LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_dual.so
for pass in 1 2 3; do
  LS_CONV_DUAL=0 timeout 300 python profiles/time_dac.py single 2>&1 | grep decode
  LS_CONV_DUAL=1 timeout 300 python profiles/time_dac.py dual 2>&1 | grep decode
done
timeout 900 python -m pytest tests -q -m gpu -x -k "dac or conv or front or enc or speaker" 2>&1 | tail -3
for d in 0 1 0 1; do
  LS_CONV_DUAL=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dual=$d step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
