#!/bin/bash
# same-box A/B/C: old commit | current | current without the fp16-operand kernels in the library
run() { (cd $1 && timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms', round(d['ms_per_step'],2), 'value', round(d['value'],1), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"); }
for rep in 1 2; do
  echo "== old"; run build_variants/wt_old
  echo "== current"; run .
  echo "== current, no fp16 kernels"; LS_LIB=$PWD/build_variants/nofp16.so LS_BUILD_DEFINES="-DLS_NO_FP16_BUILD" run .
done
