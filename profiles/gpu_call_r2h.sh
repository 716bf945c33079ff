#!/bin/bash
# same-box A/B: round-2 first commit (old kernels + new host side) vs the current tree
for d in build_variants/wt_old .; do
  echo "=== $d"
  (cd $d && timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms', round(d['ms_per_step'],2), 'value', round(d['value'],1)); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})")
done
for d in build_variants/wt_old .; do
  echo "=== $d (second pass)"
  (cd $d && timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline --no-profile 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms', round(d['ms_per_step'],2), 'value', round(d['value'],1))")
done
