#!/bin/bash
# round-1 closing call: smoke, whole GPU suite, default bench line (both arms), ncu launch list of one bench step,
# ncu --set full captures of the two dominant kernels (each only after the same command ran clean without ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; echo "bench rc=$?"; head -c 260 gpurun_out/bench_m.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_m.json 2> /dev/null; echo "ref rc=$?"
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_m.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_m.log 2>&1; echo "ncu list rc=$?"
timeout 120 python profiles/run_one.py > /dev/null 2>&1; echo "run_one rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:tblock --launch-skip 21 --launch-count 1 -o gpurun_out/r01m_tblock -f python profiles/run_one.py > gpurun_out/ncu_mt.log 2>&1; echo "ncu tblock rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn --launch-skip 20 --launch-count 1 -o gpurun_out/r01m_attn -f python profiles/run_one.py > gpurun_out/ncu_ma.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
