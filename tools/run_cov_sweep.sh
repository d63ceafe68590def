#!/usr/bin/env bash
# accumulate-kernel geometry sweep: rebuild coverage.o with -DTB_COV_TILE / -DTB_COV_RPT, parity tests once, C4-style timing each
mkdir -p gpurun_out
cd tiebrush_b200/csrc
for cfg in "4096 16" "8192 16" "8192 32"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 -DTB_COV_TILE=$1 -DTB_COV_RPT=$2 -c -o coverage.o coverage.cu 2>/dev/null
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtiebrush_b200.so capi.o coverage.o collapse.o collapse_tile.o collapse_ordered.o collapse_yd.o
  (cd ../.. && timeout 300 python bench.py --samples 4 --reads 100000 --cov-records 500000000 --cov-chroms 24 --steps 2 --warmup 1 --cpu-sample 0 --cli-reads 0 --no-e2e > gpurun_out/covsw_$1_$2.json 2> gpurun_out/covsw_$1_$2.err
   python - <<P
import json
d=json.load(open("gpurun_out/covsw_$1_$2.json"))["tiecov"]; print("tile=$1 rpt=$2", round(d["ms_per_step"],2), d["stage_ms"])
P
  )
done
