#!/bin/bash
# round-2 evidence call: smoke, whole GPU suite, default bench line (both arms), ncu launch list of one bench step,
# ncu --set full captures of tblock, attention, one estimator conv_gemm instance (resnet conv1: LN + Mish + temb) and the
# DAC conv7 at C = 48 (each only after the same command ran clean without ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; head -c 260 gpurun_out/bench_j.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_j.json 2> /dev/null; echo "ref rc=$?"
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_j.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra > gpurun_out/ncu_j.log 2>&1; echo "ncu list rc=$?"
timeout 120 python profiles/run_one.py > /dev/null 2>&1; echo "run_one rc=$?"
cap() {  # name, mangled-name regex, launch-skip
  timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base mangled \
    -k "regex:$2" --launch-skip $3 --launch-count 1 -o gpurun_out/r02j_$1 -f python profiles/run_one.py > gpurun_out/ncu_j_$1.log 2>&1
  echo "ncu $1 rc=$?"
}
cap tblock tblock_kernel 21
cap attn attn_kernel 20
cap estconv conv_gemm_kernelILi3ELi0ELi2ELi0ELi1E 4
cap dacconv7_c48 conv_gemm_kernelILi1ELi0ELi3ELi0ELi0E 12
cap dacconv1_c48 conv_gemm_kernelILi1ELi1ELi3ELi1ELi0E 10
ls -la gpurun_out/r02j_*.ncu-rep | tail -6
