#!/usr/bin/env bash
# launch list (ncu gpu__time_duration only) of one collapse step + one coverage step
TAG=${1:-ll}; READS=${2:-2000000}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'col_|yd_|tb_|cov_|junc_' -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --cpu-sample 0 --no-e2e --steps 1 --warmup 1 --samples 100 --reads $READS --cov-records 50000000 > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log | cut -c1-300
