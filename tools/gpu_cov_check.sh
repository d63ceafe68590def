#!/bin/bash
# coverage kernels after a change: parity tests of the tiecov path, then the C4 leg alone
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_coverage.py tests/test_gpu_shard.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --reads 200000 --steps 3 --warmup 2 --no-e2e --cpu-sample 0 --cli-reads 0 --cov-cpu-sample 0 $EXTRA 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())['tiecov']; print(round(d['ms_per_step'],2), d['stage_ms'], d['runs'], d['juncs'], d['roofline'])"
