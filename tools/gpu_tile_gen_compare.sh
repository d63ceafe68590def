#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/dbg_tile2.py 2>&1 | grep -v "equal" | tail -40
timeout 900 python -m pytest tests/test_gpu_collapse.py -x -q -m gpu > gpurun_out/r2b_tests_gen2.log 2>&1; echo "gen2 tests rc=$?"
tail -5 gpurun_out/r2b_tests_gen2.log
for g in 2 1; do
  TB_TILE_GEN=$g timeout 300 python bench.py --reads 2000000 --steps 5 --warmup 3 --no-e2e --cov-records 0 --cpu-sample 0 --cli-reads 0 > gpurun_out/r2b_bench_gen$g.json 2> gpurun_out/r2b_bench_gen$g.err; echo "bench gen$g rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2b_bench_gen$g.json"))
    print("gen$g", d["ms_per_step"], d["stage_ms"], d["roofline"]["frac"], d["config"].get("tile_gen"), d["config"].get("heavy_slots"), d["config"].get("tile_stats"))
except Exception as e: print("gen$g parse failed", e); print(open("gpurun_out/r2b_bench_gen$g.err").read()[-2000:])
P
done
