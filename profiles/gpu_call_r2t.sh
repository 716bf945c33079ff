#!/bin/bash
# N = 2 bench (torchrun) of the closing tree + the MMA issue-pattern micro-benchmark at more N
./profiles/micro/mma_bw | grep pattern > gpurun_out/mma_patterns_t.log; tail -24 gpurun_out/mma_patterns_t.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_t.json 2> gpurun_out/bench_n2_t.err; echo "n2 rc=$?"; head -c 300 gpurun_out/bench_n2_t.json; echo
