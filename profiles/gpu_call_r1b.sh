# PDL / halo-conv / cluster experiments: parity first, then timings of each variant (bench without the CPU baseline)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
LS_NO_PDL=1 $B > gpurun_out/bench_nopdl.json 2>&1; echo "nopdl rc=$?"
LS_CONV_HALO=0 $B > gpurun_out/bench_nohalo.json 2>&1; echo "nohalo rc=$?"
LS_LIB=$PWD/minimax-speech_b200/libls_b200_cs1.so $B > gpurun_out/bench_cs1.json 2>&1; echo "cs1 rc=$?"
LS_LIB=$PWD/minimax-speech_b200/libls_b200_cs1.so python profiles/time_kernels.py > gpurun_out/time_kernels_cs1.log 2>&1
python profiles/time_kernels.py > gpurun_out/time_kernels.log 2>&1
for f in default nopdl nohalo cs1; do python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/bench_$f.json") if l.startswith("{")][-1])
    print("$f", round(j["value"],1), "ms/step", round(j["ms_per_step"],2), {k:round(v["ms_per_step"],2) for k,v in j["kernels"].items()})
except Exception as e: print("$f", "failed", e)
PY
done
cat gpurun_out/time_kernels.log gpurun_out/time_kernels_cs1.log
