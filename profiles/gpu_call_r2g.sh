#!/bin/bash
# full GPU suite + smoke + bench (both arms) with the current build
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 300 gpurun_out/r2g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches']); print(d['roofline']); print({k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}); print(d['clocks'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_bench_ref.json 2>/dev/null; cat gpurun_out/r2g_bench_ref.json | cut -c1-400
