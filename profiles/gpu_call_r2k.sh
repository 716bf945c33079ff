#!/bin/bash
# same-box A/B: DAC residual stream fp32 (-DLS_DAC_X_F32=1) vs fp16 (default of the tree)
for pass in 1 2; do
  LS_LIB=$PWD/build_variants/libls_x32.so LS_BUILD_DEFINES="-DLS_DAC_X_F32=1" timeout 300 python profiles/time_dac.py x32 2>&1 | tail -4
  LS_LIB=$PWD/build_variants/libls_x16.so timeout 300 python profiles/time_dac.py x16 2>&1 | tail -4
done
python - <<'PY'
import torch
for init in ("reference", "trained"):
    a, b = torch.load(f"/tmp/dacwav_x32_{init}.pt"), torch.load(f"/tmp/dacwav_x16_{init}.pt")
    print(init, "fp16 vs fp32 residual: max abs diff", float((a - b).abs().max()), "SNR dB", float(10 * torch.log10(a.double().pow(2).sum() / (a - b).double().pow(2).sum())))
PY
LS_LIB=$PWD/build_variants/libls_x16.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "dac or conv" 2>&1 | tail -4
# per-tile timeline of the thin conv7 / conv1 launches (in-tree library)
LS_LIB=$PWD/build_variants/libls_x32.so LS_BUILD_DEFINES="-DLS_DAC_X_F32=1" LS_C=96 timeout 200 python profiles/timeline_dac.py 2>&1 | tail -16
LS_LIB=$PWD/build_variants/libls_x32.so LS_BUILD_DEFINES="-DLS_DAC_X_F32=1" LS_C=48 timeout 200 python profiles/timeline_dac.py 2>&1 | tail -16
