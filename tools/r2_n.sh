#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collapse.py -x -q -m gpu -k "packed" > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2n_tests.log
timeout 1200 python bench.py --steps 3 --warmup 2 --cov-records 0 --cpu-sample 0 --cli-reads 0 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2n_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench.json')); print(d['ms_per_step'], d['e2e'])"
