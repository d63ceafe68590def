#!/bin/bash
# ncu --set full of the generation-2 tile kernel (and the heavy relaunch) at 100 x 1M
mkdir -p gpurun_out
B="python bench.py --cpu-sample 0 --no-e2e --steps 1 --warmup 1 --samples 100 --reads ${READS:-1000000} --cov-records 0 --cli-reads 0"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'col_tile2_kernel|col_tile_kernel' -s 1 -c 1 -o gpurun_out/${TAG:-r2c}_prof $B > gpurun_out/${TAG:-r2c}_prof.log 2>&1
tail -3 gpurun_out/${TAG:-r2c}_prof.log | cut -c1-300
ls -la gpurun_out | grep ${TAG:-r2c}
