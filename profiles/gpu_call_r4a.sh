#!/bin/bash
# tblock: out-proj A operand (att tile) loaded straight into A3 + AH instead of through the weight ring (TBLOCK_ATT_DIRECT=1,
# the tree's default) vs the previous form (att0.so) vs ATT_DIRECT + FF2 in the TS form (ts.so): parity, isolated timing,
# in-kernel timeline, same-box A/B of the bench step
mkdir -p gpurun_out
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
for v in "default:" "att0:-DTBLOCK_ATT_DIRECT=0" "ts:-DTBLOCK_FF2_TS=1"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"
  run $n "$d" timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -2
  run $n "$d" timeout 100 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(16000,2); tk.tblock(96000,0)"
done
echo "=== timeline default"
timeout 100 python profiles/timeline_tblock.py 2>&1 | tail -24
echo "=== parity (default)"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_streaming_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -3
for n in default att0 ts default att0 ts; do
  case $n in default) d="";; att0) d="-DTBLOCK_ATT_DIRECT=0";; ts) d="-DTBLOCK_FF2_TS=1";; esac
  run $n "$d" timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
