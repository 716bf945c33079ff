#!/bin/bash
# last validation of the committed tree: smoke, whole GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"; head -c 200 gpurun_out/bench_h.json; echo
