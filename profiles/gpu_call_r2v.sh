#!/bin/bash
# N-GPU bench of the closing tree (torchrun); N from the first argument
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_v.json 2> gpurun_out/bench_n${N}_v.err; echo "n$N rc=$?"; head -c 400 gpurun_out/bench_n${N}_v.json; echo
