#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in "128 7" "128 6"; do
  set -- $cfg
  TB_TILE_THREADS=$1 TB_TILE_T8=$2 timeout 300 python bench.py --reads 2000000 --cov-records 0 --steps 3 --warmup 2 --cpu-sample 0 --no-e2e --cli-reads 0 > gpurun_out/sw_$1_$2.json 2> gpurun_out/sw_$1_$2.err
  python - <<P
import json
d=json.load(open("gpurun_out/sw_$1_$2.json")); print("thr=$1 T8=$2", "ms", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})
P
done
