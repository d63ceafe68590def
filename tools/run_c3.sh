#!/usr/bin/env bash
# C3 lines: the C2 cohort in -E and -L modes with -N/-Q/-F filters (BASELINE.json configs[2]); C4-style tiecov leg at 1e9 records
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/c3_pytest.log 2>&1; tail -3 gpurun_out/c3_pytest.log
R=${1:-4000000}
timeout 900 python bench.py --reads $R --mode 3 --max-nh 5 --min-qual 1 --cov-records 0 --steps 3 --warmup 3 --cpu-sample 0 --cli-reads 0 > gpurun_out/c3_E_NQ.json 2> gpurun_out/c3_E_NQ.err; tail -2 gpurun_out/c3_E_NQ.err; cut -c1-400 gpurun_out/c3_E_NQ.json
timeout 900 python bench.py --reads $R --mode 1 --max-nh 5 --min-qual 1 --cov-records 0 --steps 3 --warmup 3 --cpu-sample 0 --cli-reads 0 > gpurun_out/c3_L_NQ.json 2> gpurun_out/c3_L_NQ.err; tail -2 gpurun_out/c3_L_NQ.err; cut -c1-400 gpurun_out/c3_L_NQ.json
timeout 900 python bench.py --reads $R --mode 3 --flag-mask 83 --cov-records 0 --steps 3 --warmup 3 --cpu-sample 0 --cli-reads 0 > gpurun_out/c3_E_F.json 2> gpurun_out/c3_E_F.err; tail -2 gpurun_out/c3_E_F.err; cut -c1-400 gpurun_out/c3_E_F.json
python - <<P
import json
for f in ("c3_E_NQ","c3_L_NQ","c3_E_F"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"]/1e9,3),"G/s", round(d["ms_per_step"],2),"ms", {k:round(v,2) for k,v in d["stage_ms"].items()}, "path", d["config"].get("front_end_path"), "groups", d["config"]["groups_out"], "e2e", round(d["e2e"]["value"]/1e9,3))
    except Exception as e: print(f, "ERR", e)
P
