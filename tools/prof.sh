#!/usr/bin/env bash
# ncu evidence (run under gpurun, 1 GPU): launch list of OUR kernels for one bench step + one full capture of the dominant kernels.
# usage: tools/prof.sh TAG [bench args...]
set -x
mkdir -p gpurun_out
TAG=${1:-prof}; shift
B="python bench.py --cpu-sample 0 --no-e2e --steps 1 --warmup 1 ${@:---samples 100 --reads 500000 --cov-records 10000000}"
OURS='regex:col_|yd_|cov_|junc_|tb_|ord_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-col_tile_kernel|yd_chain_kernel|cov_accumulate_kernel|col_hist_kernel|col_off_kernel}" -c ${KCOUNT:-8} -o gpurun_out/${TAG}_prof $B > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out
