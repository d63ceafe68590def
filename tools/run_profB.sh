#!/usr/bin/env bash
mkdir -p gpurun_out
B="python bench.py --samples 4 --reads 100000 --cov-records 100000000 --steps 1 --warmup 0 --cpu-sample 0 --cli-reads 0 --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cov_bundle_kernel|cov_accumulate_kernel' -c 4 -o gpurun_out/profB $B > gpurun_out/profB.log 2>&1
tail -2 gpurun_out/profB.log | cut -c1-200
