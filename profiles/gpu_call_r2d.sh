#!/bin/bash
# attention v2: one CTA per item vs persistent CTAs, standalone and inside the bench step; then the full GPU test suite
for pmode in 0 1; do
  echo "=== LS_ATTN_PERSISTENT=$pmode"
  LS_ATTN_PERSISTENT=$pmode timeout 100 python -c "
import profiles.time_kernels as tk
tk.attention(32,500); tk.attention(32,500); tk.attention(64,1500)"
  LS_ATTN_PERSISTENT=$pmode timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'value', d['value']); print({k:(round(v['ms_per_step'],2)) for k,v in d['kernels'].items()})"
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
