#!/bin/bash
# tblock: sub-phase stamps of the out-proj LayerNorm epilogue (detail.so); four-chain row statistics (stats4.so)
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
echo "=== timeline detail"
run detail "-DTBLOCK_DETAIL_TL=1" timeout 100 python profiles/timeline_tblock.py 2>&1 | grep -E "^---|^EPI :|EPI out-proj"
for v in "default:" "stats4:-DTBLOCK_STATS4=1" "default:" "stats4:-DTBLOCK_STATS4=1"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"
  run $n "$d" timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -1
  run $n "$d" timeout 100 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(16000,2)"
done
