for v in "" sleep32 sleep200; do
  if [ -n "$v" ]; then export LS_LIB=$PWD/minimax-speech_b200/libls_$v.so; fi
  echo "== variant ${v:-nosleep}"; python profiles/timeline_dac.py 2>&1 | grep "==="; python profiles/time_kernels.py 2>&1 | grep -E "R=16000 tail=0|B=32"
done
