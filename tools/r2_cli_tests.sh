#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_host_cli.py tests/test_gpu_collapse.py -x -q -m gpu > gpurun_out/r2k_cli_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2k_cli_tests.log
timeout 900 python bench.py --reads 200000 --cov-records 0 --steps 2 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2k_bench_cli.json 2> gpurun_out/r2k_bench_cli.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_cli.json')); print(d.get('host_cli'))"
