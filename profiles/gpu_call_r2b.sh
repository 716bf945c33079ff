#!/bin/bash
# round 2, call b: GPU tests (new ops / graph tests first), then bench with the extra legs
python -m pytest tests/test_ops_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest_ops.log
tail -15 gpurun_out/r2b_pytest_ops.log
python -m pytest tests -m gpu -q -s --deselect tests/test_ops_gpu.py 2>&1 | grep -E "rel-L2|SNR|passed|failed|Error" | tail -60 > gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 1500 gpurun_out/r2b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['configs'], indent=1)); print(d['hbm']); print(json.dumps(d['kernels'],indent=1)); print(d['cpu_baseline'])"
