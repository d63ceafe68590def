#!/bin/bash
# K6 (cov_bundle_kernel) geometry sweep on the C4 leg: rebuilds coverage.o on the box with -DTB_CBK_ITEMS / -DTB_CBK_MINB
mkdir -p gpurun_out
cd tiebrush_b200/csrc
for cfg in "8 4" "16 2" "16 3" "8 6" "4 8"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 -DTB_CBK_ITEMS=$1 -DTB_CBK_MINB=$2 -c -o coverage.o coverage.cu 2>/dev/null && \
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtiebrush_b200.so capi.o coverage.o collapse.o collapse_tile.o collapse_ordered.o collapse_yd.o shard.o -ldl
  (cd ../.. && timeout 600 python bench.py --reads 200000 --steps 3 --warmup 2 --no-e2e --cpu-sample 0 --cli-reads 0 --cov-cpu-sample 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read())['tiecov']; print('ITEMS $1 MINB $2:', round(d['ms_per_step'],2), d['stage_ms'], d['runs'])")
done
