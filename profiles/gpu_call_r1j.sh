mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python profiles/time_encoder.py 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --batch 32 --seconds 30 --n-timesteps 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2>/dev/null; echo "cfg3 rc=$?"
LS_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:tblock --launch-skip 21 --launch-count 1 -o gpurun_out/r01_tblock -f python profiles/run_one.py > gpurun_out/ncu_t.log 2>&1; echo "ncu tblock rc=$?"
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn --launch-skip 20 --launch-count 1 -o gpurun_out/r01_attn -f python profiles/run_one.py > gpurun_out/ncu_a.log 2>&1; echo "ncu attn rc=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:conv_gemm --launch-skip 480 --launch-count 40 -o gpurun_out/r01_conv_dac -f python profiles/run_one.py > gpurun_out/ncu_c.log 2>&1; echo "ncu conv rc=$?"
