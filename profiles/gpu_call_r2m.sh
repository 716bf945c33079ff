#!/bin/bash
# weight-ring depth of the streamed-weight DAC layers: baseline (16 KB LayerNorm area always reserved, 3 A stages) vs
# v1 (area given to the rings when unused) vs v2 (v1 + 2 A stages)
for pass in 1 2; do
  LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_x16.so timeout 300 python profiles/time_dac.py base 2>&1 | grep decode
  LS_LIB=$PWD/build_variants/libls_v1.so timeout 300 python profiles/time_dac.py v1 2>&1 | grep decode
  LS_LIB=$PWD/build_variants/libls_v2.so LS_BUILD_DEFINES="-DCONV_HALO_A_STAGES=2" timeout 300 python profiles/time_dac.py v2 2>&1 | grep decode
done
for v in "x16 " "v1 " "v2 -DCONV_HALO_A_STAGES=2"; do
  set -- $v
  echo "--- $1"
  LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_$1.so LS_BUILD_DEFINES="$2" LS_C=96 timeout 200 python profiles/timeline_dac.py 2>&1 | grep -A3 "conv7 dil 9"
done
