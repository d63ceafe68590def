#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_host_cli.py -m gpu -x -q) > gpurun_out/cli_pytest.log 2>&1; tail -6 gpurun_out/cli_pytest.log
bash tools/run_cli_bench.sh
