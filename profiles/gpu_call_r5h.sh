#!/bin/bash
# closing call of the round on the FINAL tree (att-direct + TS tblock, L2 weight prefetch, reference arm on the staged
# reference modules): smoke, whole GPU suite, both bench arms, ncu launch list of one bench step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?"; head -c 200 gpurun_out/bench_i.json; echo
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_i.json 2> /dev/null; echo "ref rc=$?"; head -c 120 gpurun_out/bench_ref_i.json; echo
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_i.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra > gpurun_out/ncu_i.log 2>&1; echo "ncu list rc=$?"
