#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python bench.py --samples 1000 --reads 100000 --cov-records 0 --steps 3 --warmup 2 --cpu-sample 2000000 --cli-reads 0 > gpurun_out/k1000.json 2> gpurun_out/k1000.err; tail -3 gpurun_out/k1000.err
python - <<P
import json
d=json.load(open("gpurun_out/k1000.json")); print("k=1000", round(d["value"]/1e9,3), "G/s ms", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, "groups", d["config"]["groups_out"], "e2e", d.get("e2e",{}).get("value"), d.get("cpu_baseline"))
P
