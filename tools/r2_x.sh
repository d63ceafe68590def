#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_host_cli.py -x -q -m gpu -k "sample or frac or heat or recollapse or exact" > gpurun_out/r2x_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2x_tests.log
timeout 600 python tools/run_sample_time.py 2>&1 | tail -3
TB_COV_WALK=brute timeout 600 python tools/run_sample_time.py 2>&1 | tail -2
