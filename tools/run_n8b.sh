#!/usr/bin/env bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --reads 4000000 --steps 2 --warmup 3 --cov-records 0 --cpu-sample 0 --cli-reads 0) > gpurun_out/n${N}b_bench.json 2> gpurun_out/n${N}b_bench.err; tail -3 gpurun_out/n${N}b_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/n${N}b_bench.json").read().strip().split("\n")[-1]); print("N=$N value", d["value"]/1e9, "e2e", d["e2e"], d["config"].get("host_affinity"))
P
