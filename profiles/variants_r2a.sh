#!/bin/bash
# round 2: (1) how does the fused block kernel's time depend on the weight ring depth?  (2) Snake with / without range reduction
run() { # name defines script...
  local name=$1 defs=$2; shift 2
  if [ "$name" = default ]; then "$@"; else LS_LIB=$PWD/build_variants/$name.so LS_BUILD_DEFINES="$defs" "$@"; fi
}
for v in "default:" "slots4:-DTBLOCK_SLOTS=4" "slots3:-DTBLOCK_SLOTS=3"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"; run $n "$d" python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(96000,0)"
done
for v in "default:" "snake0:-DLS_SNAKE_REDUCE=0"; do
  n=${v%%:*}; d=${v#*:}
  echo "=== $n"; run $n "$d" python -m pytest tests/test_parity_gpu.py -q -s -k "trained or noncausal or dac_decode_vs" 2>&1 | grep -E "SNR|rel-L2|passed|failed"
  run $n "$d" python -c "
import torch, minimax_speech_b200.synth as synth, profiles.time_kernels as tk
from minimax_speech_b200.dac import DACVAEDecoder
dec = DACVAEDecoder(); dec.load_state_dict(synth.dac_decoder_state_dict(0,'reference'))
z = torch.randn(16,80,500,device='cuda')
print('dac decode 16x10s: %.1f us' % tk.timeit(lambda: dec.decode(z)))"
done
