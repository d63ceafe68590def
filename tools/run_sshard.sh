#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 200 python -m pytest tests/test_gpu_shard.py -m gpu -x -q) > gpurun_out/sshard_pytest.log 2>&1; tail -6 gpurun_out/sshard_pytest.log
