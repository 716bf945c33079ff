#!/bin/bash
# tblock: FF2 in TS mode (default) and the second weight ring (TBLOCK_RING_B=1 build)
for v in "ringb4:-DTBLOCK_RING_B=1 -DTBLOCK_PRODUCER_WARPS=4"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"
  if [ "$n" = default ]; then E=""; else export LS_LIB=$PWD/build_variants/$n.so LS_BUILD_DEFINES="$d"; fi
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -3
  timeout 120 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(96000,0)"
  timeout 120 python profiles/timeline_tblock.py 2>&1 | grep -v "^LOAD\|^MMA saw" | sed -n 4,6p
  unset LS_LIB LS_BUILD_DEFINES
done
