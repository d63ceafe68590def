#!/usr/bin/env bash
# slot-size experiment for the tile kernel
mkdir -p gpurun_out
for t8 in 4 5 6 7 8; do
  TB_TILE_T8=$t8 timeout 300 python bench.py --reads 2000000 --cov-records 0 --steps 3 --warmup 2 --cpu-sample 0 --no-e2e > gpurun_out/t8_$t8.json 2> gpurun_out/t8_$t8.err
  python - <<P
import json
d=json.load(open("gpurun_out/t8_$t8.json")); print("T8=$t8", "groups", d["config"]["groups_out"], "ms", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})
P
done
for thr in 512 1024; do
  TB_TILE_T8=7 TB_TILE_THREADS=$thr timeout 300 python bench.py --reads 2000000 --cov-records 0 --steps 3 --warmup 2 --cpu-sample 0 --no-e2e > gpurun_out/t8_thr$thr.json 2> gpurun_out/t8_thr$thr.err
  python - <<P
import json
d=json.load(open("gpurun_out/t8_thr$thr.json")); print("T8=7 thr=$thr", "groups", d["config"]["groups_out"], "ms", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})
P
done
