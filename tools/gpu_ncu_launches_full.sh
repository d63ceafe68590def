#!/bin/bash
# ncu launch list (one metric, one pass) of one full-size step: every kernel's share of the step
mkdir -p gpurun_out
OURS='regex:col_|yd_|cov_|junc_|tb_|ord_|shard_'
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 2000 --csv --log-file gpurun_out/${TAG:-r2r}_launches_full.csv python bench.py --steps 1 --warmup 1 --no-e2e --cpu-sample 0 --cli-reads 0 --cov-cpu-sample 0 ${EXTRA} > gpurun_out/${TAG:-r2r}_launches_full.log 2>&1
echo rc=$?; python tools/ncu_launch_summary.py gpurun_out/${TAG:-r2r}_launches_full.csv | tail -75
