#!/bin/bash
# tiecov C4 leg of bench.py at a reduced size, 1 GPU and (if present) all GPUs under torchrun
mkdir -p gpurun_out
R=${R:-200000000}
NG=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 900 python bench.py --reads 200000 --cov-records $R --cov-window 12500000 --cov-e2e-records 50000000 --steps 3 --warmup 2 --cpu-sample 0 --cli-reads 0 > gpurun_out/${TAG:-r2g}_n1.json 2> gpurun_out/${TAG:-r2g}_n1.err; echo "n1 rc=$?"; tail -3 gpurun_out/${TAG:-r2g}_n1.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG:-r2g}_n1.json"))["tiecov"]
print(json.dumps({k:d[k] for k in d if k!="config"}, indent=None)[:2500])
P
if [ "$NG" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $NG --reads 200000 --cov-records $R --cov-window 12500000 --cov-e2e-records 50000000 --steps 3 --warmup 2 --cpu-sample 0 --cli-reads 0 > gpurun_out/${TAG:-r2g}_n$NG.json 2> gpurun_out/${TAG:-r2g}_n$NG.err; echo "n$NG rc=$?"; tail -3 gpurun_out/${TAG:-r2g}_n$NG.err
  python - <<P
import json
d=json.loads(open("gpurun_out/${TAG:-r2g}_n$NG.json").read().strip().splitlines()[-1])["tiecov"]
print(json.dumps({k:d[k] for k in d if k!="config"}, indent=None)[:2500])
P
fi
