#!/bin/bash
# coverage.cu compile-time geometry sweep on the C4 leg: CFGS is a ';'-separated list of -D flag sets; coverage.o is rebuilt
# on the box for each (the local build is not touched)
mkdir -p gpurun_out
cd tiebrush_b200/csrc
IFS=';' read -ra LIST <<< "${CFGS:--DTB_COV_BATCH=4}"
for cfg in "${LIST[@]}"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 $cfg -c -o coverage.o coverage.cu 2>/dev/null && \
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtiebrush_b200.so capi.o coverage.o collapse.o collapse_tile.o collapse_ordered.o collapse_yd.o shard.o -ldl
  (cd ../.. && timeout 600 python bench.py --reads 200000 --steps 3 --warmup 2 --no-e2e --cpu-sample 0 --cli-reads 0 --cov-cpu-sample 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())['tiecov']; print('$cfg:', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stage_ms'].items()}, d['runs'], d['juncs'])")
done
