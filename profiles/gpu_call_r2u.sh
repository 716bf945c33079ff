#!/bin/bash
# S3 tokenizer trunk (fp32 mode): parity tests + a timing of the shipped configuration
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -s -k "s3 or fsq" 2>&1 | tail -12
timeout 300 python - <<'PY'
import torch, time
import minimax_speech_b200.synth as synth
from minimax_speech_b200.tokenizer import S3TokenizerV2
tok = S3TokenizerV2(weight_seed=33)
dev = torch.device("cuda:0")
for B, T in ((1, 1000), (16, 1000)):
    mel = torch.cat([synth.s3_mel(i, T) for i in range(B)], 0).to(dev); ml = torch.full((B,), T, dtype=torch.int32, device=dev)
    tok.quantize(mel, ml); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tok.quantize(mel, ml); e1.record(); torch.cuda.synchronize()
    print(f"s3 quantize {B} x {T / 100:.0f} s: {e0.elapsed_time(e1):.1f} ms")
PY
