mkdir -p gpurun_out
echo "== halo mode 2 (default)"; python -m pytest tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -8
echo "== halo mode 1"; LS_CONV_HALO=1 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv or linear or gelu or k_split or lengths" 2>&1 | tail -5
echo "== halo mode 0"; LS_CONV_HALO=0 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv or linear or gelu or k_split or lengths" 2>&1 | tail -3
python profiles/time_kernels.py 2>&1 | tail -8
