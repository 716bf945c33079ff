#!/bin/bash
# round 2, call a: GPU tests of the new call surface + bench with the extra legs
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 600 gpurun_out/r2a_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2a_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['configs'], indent=1)); print(d['hbm']); print(d['kernels'])"
