#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_shard.py -m gpu -x -q) > gpurun_out/cov3_pytest.log 2>&1; tail -5 gpurun_out/cov3_pytest.log
timeout 300 python bench.py --samples 4 --reads 100000 --cov-records 100000000 --steps 3 --warmup 3 --cpu-sample 0 --cli-reads 0 --no-e2e > gpurun_out/cov3_bench.json 2> gpurun_out/cov3_bench.err; tail -3 gpurun_out/cov3_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/cov3_bench.json")); print(json.dumps(d["tiecov"]))
P
bash tools/run_c4.sh
