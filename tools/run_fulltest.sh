#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_collapse.py tests/test_gpu_coverage.py -m gpu -x -q -k "baseline_size or c4_scale") > gpurun_out/fulltest_pytest.log 2>&1; tail -25 gpurun_out/fulltest_pytest.log
