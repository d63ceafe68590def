#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_collapse.py -x -q -m gpu -k "packed or c3_full or wire" > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2m_tests.log
TAG=r2m bash tools/r2_bench_default.sh
