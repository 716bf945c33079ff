#!/bin/bash
# tblock changes: kernel parity tests, timing, timeline
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tblock" 2>&1 | tail -4
timeout 120 python -c "
import profiles.time_kernels as tk
tk.tblock(16000,0); tk.tblock(16000,0); tk.tblock(16000,1); tk.tblock(16000,2); tk.tblock(96000,0)"
timeout 120 python profiles/timeline_tblock.py 2>&1 | grep -v "^LOAD\|^MMA saw" | sed -n 3,11p
