#!/usr/bin/env bash
# r01f: full-size C2 bench line (default flags), reference arm, then ncu --set full of the dominant kernels at 100 x 2M
mkdir -p gpurun_out
(time timeout 1200 python bench.py --steps 3 --warmup 3) > gpurun_out/r01f_bench_C2_full.json 2> gpurun_out/r01f_bench_C2_full.err; tail -4 gpurun_out/r01f_bench_C2_full.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r01f_reference_arm.json 2> gpurun_out/r01f_reference_arm.err; tail -4 gpurun_out/r01f_reference_arm.err
B="python bench.py --cpu-sample 0 --no-e2e --steps 1 --warmup 1 --samples 100 --reads 2000000 --cov-records 50000000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'col_tile_kernel|yd_frontier_kernel|cov_accumulate_kernel|yd_desc_kernel|yd_lookup_kernel|col_hist_kernel' -c 12 -o gpurun_out/r01f_prof $B > gpurun_out/r01f_prof.log 2>&1
tail -2 gpurun_out/r01f_prof.log | cut -c1-200; ls -la gpurun_out | tail -5
