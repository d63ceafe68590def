#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coverage.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2q_tests.log
for extra in "" "--cov-no-end"; do
timeout 900 python bench.py --reads 200000 --steps 3 --warmup 2 --no-e2e --cpu-sample 0 --cli-reads 0 --cov-cpu-sample 0 $extra > gpurun_out/r2q_bench$extra.json 2> gpurun_out/r2q_bench$extra.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench$extra.err
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench$extra.json'))['tiecov']; print('$extra', d['ms_per_step'], d['records_per_sec'], d['stage_ms'], d['config']['windows_per_gpu'], d['runs'], d['juncs'], d['roofline']['frac'])"
done
