#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_collapse.py -m gpu -x -q -k "compact or baseline_size") > gpurun_out/wire_pytest.log 2>&1; tail -4 gpurun_out/wire_pytest.log
(time timeout 1500 python bench.py --steps 3 --warmup 3 --cli-reads 0) > gpurun_out/wire_bench.json 2> gpurun_out/wire_bench.err; tail -4 gpurun_out/wire_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/wire_bench.json")); print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"], "cpu", d.get("cpu_baseline"))
P
