#!/bin/bash
# tblock: out-proj LayerNorm epilogue keeps columns 32..63 in registers across the statistics barrier (default) vs the previous library
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; elif [ "$name" = prev ]; then LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/prev.so "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
echo "=== timeline detail"
run detail "-DTBLOCK_DETAIL_TL=1" timeout 100 python profiles/timeline_tblock.py 2>&1 | grep -E "^---|^EPI :|EPI out-proj"
for n in default prev default prev; do
  echo "=== $n"
  run $n "" timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -1
  run $n "" timeout 100 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(16000,2)"
done
