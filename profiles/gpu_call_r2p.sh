#!/bin/bash
# dual-issue mode of conv_gemm: parity tests, then same-library A/B (LS_CONV_DUAL=0 / 1) of the DAC decode and timelines
export LS_NO_REBUILD=1 LS_LIB=$PWD/build_variants/libls_dual.so
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "dac or conv" 2>&1 | tail -4
for pass in 1 2; do
  LS_CONV_DUAL=0 timeout 300 python profiles/time_dac.py single 2>&1 | grep decode
  LS_CONV_DUAL=1 timeout 300 python profiles/time_dac.py dual 2>&1 | grep decode
done
for c in 96 48; do
  for d in 0 1; do
    echo "--- C=$c dual=$d"
    LS_CONV_DUAL=$d LS_C=$c timeout 200 python profiles/timeline_dac.py 2>&1 | grep -A3 "conv7 dil"
  done
done
