#!/bin/bash
# attention v2: share of the exponentials computed on the FMA pipe (polynomial) instead of MUFU
run() { local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi; }
for v in "default:" "poly11:-DATTN_POLY_MASK=0x11" "poly55:-DATTN_POLY_MASK=0x55" "poly77:-DATTN_POLY_MASK=0x77"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"
  run $n "$d" timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -2
  run $n "$d" timeout 100 python -c "
import profiles.time_kernels as tk
tk.attention(32,500); tk.attention(32,500); tk.attention(64,1500)"
done
