#!/bin/bash
# round 2: first run of the generation-2 tile kernel: collapse GPU tests (both generations), then a mid-size bench line per generation
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collapse.py -x -q -m gpu > gpurun_out/r2a_tests_gen2.log 2>&1; echo "gen2 tests rc=$?" | tee -a gpurun_out/r2a_tests_gen2.log
tail -5 gpurun_out/r2a_tests_gen2.log
for g in 2 1; do
  TB_TILE_GEN=$g timeout 300 python bench.py --reads 2000000 --steps 5 --warmup 3 --no-e2e --cov-records 0 --cpu-sample 0 --cli-reads 0 > gpurun_out/r2a_bench_gen$g.json 2> gpurun_out/r2a_bench_gen$g.err; echo "bench gen$g rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/r2a_bench_gen$g.json"))
    print("gen$g", d["ms_per_step"], d["stage_ms"], d["roofline"]["frac"], d["config"].get("tile_gen"), d["config"].get("heavy_slots"), d["config"].get("tile_stats"))
except Exception as e: print("gen$g parse failed", e); print(open("gpurun_out/r2a_bench_gen$g.err").read()[-2000:])
P
done
