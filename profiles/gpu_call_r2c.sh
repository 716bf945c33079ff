#!/bin/bash
# attention v2 (TS-mode MMAs): kernel parity tests, timing, timeline
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -15
timeout 120 python -c "
import profiles.time_kernels as tk
tk.attention(32,500); tk.attention(32,500); tk.attention(64,1500); tk.attention(2,500)"
timeout 120 python profiles/timeline_attn.py 2>&1 | sed -n 3,9p
