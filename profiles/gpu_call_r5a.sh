#!/bin/bash
# closing call of the round, final tree (tblock: att tile direct + TS form; reference arm on the staged reference modules):
# smoke, whole GPU suite, both bench arms, ncu launch list of one bench step, --set full of one full-mode tblock launch
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; head -c 200 gpurun_out/bench_g.json; echo; tail -3 gpurun_out/bench_g.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; echo "ref rc=$?"; head -c 120 gpurun_out/bench_ref_g.json; echo; tail -2 gpurun_out/bench_ref_g.err
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_g.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra > gpurun_out/ncu_g.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off \
  -k "regex:tblock_kernel" --launch-skip 2 --launch-count 1 -o gpurun_out/r02g_tblock -f python profiles/run_one.py > gpurun_out/ncu_g_tblock.log 2>&1; echo "ncu tblock rc=$?"
ls -la gpurun_out/r02g_*.ncu-rep | tail -2
