#!/usr/bin/env bash
# ncu evidence for round 1: launch list + one full capture of the dominant kernels (run under gpurun, 1 GPU).
set -x
mkdir -p gpurun_out
TAG=${1:-r01a}
B="python bench.py --samples 100 --reads 500000 --cov-records 10000000 --cpu-sample 0 --no-e2e --steps 2 --warmup 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-col_tile_kernel|yd_chain_kernel|cov_accumulate_kernel|col_hist_kernel}" -c ${KCOUNT:-6} -o gpurun_out/${TAG}_prof $B > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out
