#!/bin/bash
# who waits for whom in the streamed-weight conv7 at C = 96: per-tap stamps of the MMA warp (d1), per-load stamps of weight
# producer 1 (d2); 4 / 5 producer warps
for v in "d1 -DCONV_DETAIL_TL=1" "d2 -DCONV_DETAIL_TL=2" "p4 -DCONV_PRODUCER_WARPS=4" "p5 -DCONV_PRODUCER_WARPS=5"; do
  set -- $v
  echo "--- $1"
  LS_DETAIL=1 LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_$1.so LS_BUILD_DEFINES="$2" LS_C=96 timeout 200 python profiles/timeline_dac.py 2>&1 | grep -A5 "conv7 dil 9" | grep -v "MMA saw"
done
for v in "p4 -DCONV_PRODUCER_WARPS=4" "p5 -DCONV_PRODUCER_WARPS=5" "v1 "; do
  set -- $v
  LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_$1.so LS_BUILD_DEFINES="$2" timeout 300 python profiles/time_dac.py $1 2>&1 | grep decode
done
