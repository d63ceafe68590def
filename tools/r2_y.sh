#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_collapse.py tests/test_gpu_host_cli.py tests/test_gpu_shard.py -x -q -m gpu -k "not full_size and not baseline_size" > gpurun_out/r2y_tests.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2y_tests.log
for deep in 96 0; do
TB_ORD_DEEP=$deep timeout 900 python bench.py --mode 3 --flag-mask 83 --reads 4000000 --steps 2 --warmup 1 --no-e2e --cov-records 0 --cpu-sample 0 --cli-reads 0 > gpurun_out/r2y_c3_F_deep$deep.json 2> gpurun_out/r2y_c3_F_deep$deep.err; echo "bench rc=$?"; tail -2 gpurun_out/r2y_c3_F_deep$deep.err
python -c "
import json; d=json.load(open('gpurun_out/r2y_c3_F_deep$deep.json')); print('deep $deep', d['value'], d['ms_per_step'], d['config']['groups_out'], d['config']['front_end_path'])"
done
