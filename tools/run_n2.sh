#!/usr/bin/env bash
# 2-GPU check: bench line under torchrun (weak scaling, one shard per rank) + the NCCL halo / ordered-gather check
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --reads 4000000 --steps 3 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; tail -3 gpurun_out/n2_bench.err; cut -c1-600 gpurun_out/n2_bench.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_nccl_check.py > gpurun_out/n2_nccl.log 2>&1; tail -5 gpurun_out/n2_nccl.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/n2_ref.json 2> gpurun_out/n2_ref.err; cut -c1-300 gpurun_out/n2_ref.json
