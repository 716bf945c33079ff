#!/bin/bash
# round-2 closing call, final tree (fp16 DAC residual, dual-issue conv, fused res_conv, S3 tokenizer): smoke, whole GPU suite,
# both bench arms, ncu launch list of one bench step, --set full of the fused resnet conv2 launch
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; head -c 200 gpurun_out/bench_f.json; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_f.json 2> /dev/null; echo "ref rc=$?"
timeout 600 env LS_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-extra > gpurun_out/ncu_f.log 2>&1; echo "ncu list rc=$?"
timeout 120 python profiles/run_one.py > /dev/null 2>&1; echo "run_one rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base mangled \
  -k "regex:conv_gemm_kernelILi3ELi1ELi0ELi4ELi0E" --launch-skip 4 --launch-count 1 -o gpurun_out/r02f_estconv2_fused -f python profiles/run_one.py > gpurun_out/ncu_f_conv2.log 2>&1; echo "ncu conv2 rc=$?"
ls -la gpurun_out/r02f_*.ncu-rep | tail -2
