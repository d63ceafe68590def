#!/usr/bin/env bash
# r01k: round-1 closing measurements: full GPU test suite, full-size C2 bench line (default flags), reference arm, launch list at 100 x 2M
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r01k_pytest.log 2>&1; tail -4 gpurun_out/r01k_pytest.log
(time timeout 1500 python bench.py --steps 3 --warmup 3) > gpurun_out/r01k_bench_C2_full.json 2> gpurun_out/r01k_bench_C2_full.err; tail -4 gpurun_out/r01k_bench_C2_full.err
(time timeout 900 python bench.py --impl reference --steps 2 --warmup 1) > gpurun_out/r01k_reference_arm.json 2> gpurun_out/r01k_reference_arm.err; tail -4 gpurun_out/r01k_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'col_|yd_|tb_|cov_|junc_' -c 800 --csv --log-file gpurun_out/r01k_launches.csv python bench.py --cpu-sample 0 --no-e2e --cli-reads 0 --steps 1 --warmup 1 --samples 100 --reads 2000000 --cov-records 50000000 > gpurun_out/r01k_launches.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
