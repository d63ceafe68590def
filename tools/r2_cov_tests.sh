#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/r2f_cov_tests.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2f_cov_tests.log
