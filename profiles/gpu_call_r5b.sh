#!/bin/bash
# closing N = 2 line of the final tree (torchrun, NCCL, the configs[2..4] legs included), as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_g.json 2> gpurun_out/bench_n2_g.err; echo "n2 rc=$?"
head -c 300 gpurun_out/bench_n2_g.json; echo; tail -3 gpurun_out/bench_n2_g.err
