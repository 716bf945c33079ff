#!/bin/bash
# round-2 closing call (tree with the fp16 DAC residual stream and the dual-issue conv mode): smoke, whole GPU suite,
# both bench arms, ncu launch list of one bench step, --set full of the DAC conv1 / conv7 at C = 48 and C = 96
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err; echo "bench rc=$?"; head -c 200 gpurun_out/bench_s.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s.json 2> /dev/null; echo "ref rc=$?"
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_s.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra > gpurun_out/ncu_s.log 2>&1; echo "ncu list rc=$?"
timeout 120 python profiles/run_one.py > /dev/null 2>&1; echo "run_one rc=$?"
cap() {  # name, mangled-name regex, launch-skip
  timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base mangled \
    -k "regex:$2" --launch-skip $3 --launch-count 1 -o gpurun_out/r02s_$1 -f python profiles/run_one.py > gpurun_out/ncu_s_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap dacconv7_c48 conv_gemm_kernelILi1ELi0ELi3ELi0ELi0E 15
cap dacconv7_c96 conv_gemm_kernelILi1ELi0ELi3ELi0ELi0E 12
cap dacconv1_c48 conv_gemm_kernelILi1ELi3ELi3ELi3ELi0E 8
ls -la gpurun_out/r02s_*.ncu-rep | tail -4
