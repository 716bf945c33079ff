#!/bin/bash
# tblock: nine-slot weight ring in the FF phase (TBLOCK_WIDE_FF=1, the tree's default: AH boxes join the ring, TS form of FF2,
# table-driven slots with even / odd release barriers) vs ATT_DIRECT + TS with the five-slot ring (ts.so)
mkdir -p gpurun_out
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
TS="-DTBLOCK_WIDE_FF=0 -DTBLOCK_FF2_TS=1"
for v in "default:" "ts:$TS"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"
  run $n "$d" timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -2
  run $n "$d" timeout 100 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(16000,2); tk.tblock(96000,0)"
done
echo "=== timeline default"
timeout 100 python profiles/timeline_tblock.py 2>&1 | grep -E "CTA|MMA :|EPI :|LOAD|saw" | head -12
echo "=== parity (default)"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_streaming_gpu.py tests/test_boundary_gpu.py -q -m gpu -x -k "not s3" 2>&1 | tail -3
for n in default ts default ts; do
  case $n in default) d="";; ts) d="$TS";; esac
  run $n "$d" timeout 300 python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n step ms', round(d['ms_per_step'],2), {k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
