#!/usr/bin/env bash
# coverage-only check: parity tests of the coverage path + the tiecov bench leg (TAG = output prefix)
TAG=${1:-cov}; REC=${2:-100000000}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_shard.py -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --samples 20 --reads 200000 --cov-records $REC --steps 3 --warmup 3 --cpu-sample 0 --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(json.dumps(d["tiecov"]))
P
