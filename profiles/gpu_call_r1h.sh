mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -c 600 gpurun_out/bench_ref.json
LS_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc=$?"
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:tblock --launch-skip 20 --launch-count 2 -o gpurun_out/r01_tblock -f python profiles/run_one.py > gpurun_out/ncu_t.log 2>&1; echo "ncu tblock rc=$?"
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attn --launch-skip 20 --launch-count 1 -o gpurun_out/r01_attn -f python profiles/run_one.py > gpurun_out/ncu_a.log 2>&1; echo "ncu attn rc=$?"
