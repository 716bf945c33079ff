python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention 2>&1 | tail -2
for v in "" 0x00 0x11 0x49; do
  if [ -n "$v" ]; then export LS_LIB=$PWD/minimax-speech_b200/libls_attn_$v.so; fi
  echo "variant ${v:-default(0x55)}"; python profiles/time_kernels.py 2>&1 | grep attention
done
