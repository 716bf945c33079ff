for v in "" 0x11 0x49 0x55; do
  if [ -n "$v" ]; then export LS_LIB=$PWD/minimax-speech_b200/libls_poly$v.so; fi
  echo "variant poly mask ${v:-default(0x00)}"; python profiles/time_kernels.py 2>&1 | grep attention
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention 2>&1 | tail -1
done
