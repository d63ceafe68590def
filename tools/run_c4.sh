#!/usr/bin/env bash
# C4-style tiecov leg: whole-genome collapsed stream, 5e8 records per window (C4 = 2e9 records = 4 such windows per GPU)
mkdir -p gpurun_out
timeout 900 python bench.py --samples 4 --reads 100000 --cov-records ${1:-500000000} --cov-chroms 24 --steps 2 --warmup 1 --cpu-sample 0 --cli-reads 0 --no-e2e > gpurun_out/c4_cov.json 2> gpurun_out/c4_cov.err; tail -2 gpurun_out/c4_cov.err
python - <<P
import json
d=json.load(open("gpurun_out/c4_cov.json")); print("c4", json.dumps(d["tiecov"]))
P
