#!/bin/bash
# full-size C2 (100 x 10M): generation 2 vs generation 1 tile kernel, device-resident leg only
mkdir -p gpurun_out
for g in ${GENS:-2 1}; do
  TB_TILE_GEN=$g timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --cov-records 0 --cpu-sample 0 --cli-reads 0 > gpurun_out/${TAG:-r2d}_bench_full_gen$g.json 2> gpurun_out/${TAG:-r2d}_bench_full_gen$g.err; echo "bench gen$g rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/${TAG:-r2d}_bench_full_gen$g.json"))
    print("gen$g", d["ms_per_step"], d["stage_ms"], d["roofline"]["frac"], d["config"].get("tile_gen"), d["config"].get("heavy_slots"), d["config"].get("tile_stats"), d["config"]["groups_out"])
except Exception as e: print("gen$g parse failed", e); print(open("gpurun_out/${TAG:-r2d}_bench_full_gen$g.err").read()[-2000:])
P
done
