#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py --reads 200000 --cov-records 0 --steps 2 --warmup 1 --cpu-sample 0 --no-e2e --cli-reads 50000 > gpurun_out/clib.json 2> gpurun_out/clib.err; tail -3 gpurun_out/clib.err
python - <<P
import json
d=json.load(open("gpurun_out/clib.json")); print(json.dumps(d["host_cli"]))
P
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 --ref-reads 50000 | cut -c1-400
