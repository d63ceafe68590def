#!/usr/bin/env bash
# final line of the round with default flags (what the driver runs), plus smoke
mkdir -p gpurun_out
(time timeout 1500 python bench.py) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/final_bench.json")); print("value", d["value"]/1e9, "ms", d["ms_per_step"], d["stage_ms"], "roof", d["roofline"]["frac"]); print("e2e", d["e2e"]); print("tiecov", d["tiecov"]["value"]/1e12, d["tiecov"]["stage_ms"]); print("cli", d["host_cli"]); print("cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"], d["clocks"])
P
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
