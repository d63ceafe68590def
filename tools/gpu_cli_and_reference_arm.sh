#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_cli.py -x -q -m gpu > gpurun_out/r2l_cli_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2l_cli_tests.log
timeout 900 python bench.py --reads 200000 --cov-records 0 --steps 2 --warmup 1 --no-e2e --cpu-sample 0 > gpurun_out/r2l_bench_cli.json 2> gpurun_out/r2l_bench_cli.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_cli.json')); print(d.get('host_cli'))"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_reference_arm.json 2>gpurun_out/r2l_reference_arm.err; cut -c1-1500 gpurun_out/r2l_reference_arm.json
