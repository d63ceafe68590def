#!/bin/bash
# the driver's launch line on all GPUs of the box: full default sizes
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
(time timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $NG --steps ${STEPS:-5} --warmup 3 $EXTRA) > gpurun_out/${TAG:-r2i}_bench_n$NG.json 2> gpurun_out/${TAG:-r2i}_bench_n$NG.err; echo "rc=$?"; tail -5 gpurun_out/${TAG:-r2i}_bench_n$NG.err
python - <<P
import json
d=json.loads(open("gpurun_out/${TAG:-r2i}_bench_n$NG.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"stage",d["stage_ms"],"e2e",d.get("e2e",{}))
t=d["tiecov"]; print("tiecov", {k:t[k] for k in t if k not in ("config","roofline")}); print(t["config"])
P
