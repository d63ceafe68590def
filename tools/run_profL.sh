#!/usr/bin/env bash
mkdir -p gpurun_out
B="python bench.py --cpu-sample 0 --no-e2e --cli-reads 0 --steps 1 --warmup 0 --samples 100 --reads 1000000 --cov-records 0 --mode 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'col_tile_kernel' -c 1 -o gpurun_out/profL $B > gpurun_out/profL.log 2>&1
tail -2 gpurun_out/profL.log | cut -c1-200
