#!/bin/bash
# round-1 late call: whole GPU suite, default bench line (both arms), ncu launch list of one bench step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_k.log; cat gpurun_out/pytest_k.log
timeout 600 python bench.py > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -c 600 gpurun_out/bench_k.err; head -c 1500 gpurun_out/bench_k.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_k.json 2> gpurun_out/bench_ref_k.err; cat gpurun_out/bench_ref_k.json
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_k.log 2>&1
tail -3 gpurun_out/ncu_k.log; wc -l gpurun_out/launches_k.csv
