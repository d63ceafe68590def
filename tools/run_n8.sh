#!/usr/bin/env bash
# 8-GPU check with the driver's own launch line and the default (full-size) configuration
mkdir -p gpurun_out
N=${1:-8}
(time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3) > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; tail -4 gpurun_out/n${N}_bench.err; tail -1 gpurun_out/n${N}_bench.json | cut -c1-300
free -g | head -2; nproc
