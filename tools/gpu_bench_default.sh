#!/bin/bash
# the driver's default line: python bench.py (N = 1), then the reference arm
mkdir -p gpurun_out
(time timeout 1500 python bench.py --steps ${STEPS:-5} --warmup 3) > gpurun_out/${TAG:-r2h}_bench_default.json 2> gpurun_out/${TAG:-r2h}_bench_default.err; echo "rc=$?"; tail -5 gpurun_out/${TAG:-r2h}_bench_default.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG:-r2h}_bench_default.json"))
print("value",d["value"],"ms",d["ms_per_step"],"stage",d["stage_ms"],"roof",d["roofline"]["frac"],"e2e",d.get("e2e",{}))
t=d["tiecov"]; print("tiecov", {k:t[k] for k in t if k not in ("config","roofline")}); print(t["config"]); print(t["roofline"])
print("cpu", d.get("cpu_baseline")); print("cli", d.get("host_cli"))
P
