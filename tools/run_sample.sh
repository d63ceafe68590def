#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests/test_gpu_coverage.py -m gpu -x -q -k "sample_heatmap or golden") > gpurun_out/sample_pytest.log 2>&1; tail -25 gpurun_out/sample_pytest.log
