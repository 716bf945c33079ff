mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python profiles/conv1_variants.py 2>&1 | tail -8
LS_C=48 python profiles/conv1_variants.py 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/bench_j.json") if l.startswith("{")][-1])
print(round(j["value"],1), "audio-s/s; ms/step", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), {k:round(v["ms_per_step"],2) for k,v in j["kernels"].items()})
PY
