#!/usr/bin/env bash
# quick GPU check: parity tests + one mid-size bench line (TAG = output prefix)
TAG=${1:-q}; READS=${2:-2000000}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --reads $READS --cov-records 50000000 --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value",d["value"],"ms",d["ms_per_step"],"stage",d["stage_ms"],"roof",d["roofline"]["frac"],"e2e",d.get("e2e",{}).get("value"))
P
