for v in "" 4 8; do
  if [ -n "$v" ]; then export LS_LIB=$PWD/minimax-speech_b200/libls_wl$v.so; fi
  echo "== weight lanes ${v:-0 (warps)}"
  python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -1
  python profiles/conv1_variants.py 2>&1 | tail -3
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g$v.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/bench_g$v.json") if l.startswith("{")][-1])
print(round(j["value"],1), "audio-s/s; ms/step", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["value"],1), {k:round(v["ms_per_step"],2) for k,v in j["kernels"].items()})
PY
done
