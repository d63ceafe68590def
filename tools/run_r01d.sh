#!/usr/bin/env bash
# r01d: sanity of the GPU tests on this build, one mid-size bench line, launch list + full ncu of the dominant kernels
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r01d_pytest.log 2>&1; tail -3 gpurun_out/r01d_pytest.log
timeout 600 python bench.py --reads 2000000 --cov-records 50000000 --steps 3 --warmup 3 > gpurun_out/r01d_bench_2M.json 2> gpurun_out/r01d_bench_2M.err; tail -2 gpurun_out/r01d_bench_2M.err; cat gpurun_out/r01d_bench_2M.json
B="python bench.py --cpu-sample 0 --no-e2e --steps 1 --warmup 1 --samples 100 --reads 1000000 --cov-records 10000000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'col_off_kernel|col_tile_kernel|yd_desc_kernel|yd_frontier_kernel|yd_link_kernel|yd_scatter_kernel|cov_accumulate_kernel' -c 7 -o gpurun_out/r01d_prof $B > gpurun_out/r01d_prof.log 2>&1
tail -3 gpurun_out/r01d_prof.log; ls -la gpurun_out
