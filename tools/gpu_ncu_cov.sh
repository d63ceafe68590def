#!/bin/bash
# ncu --set full of the coverage kernels (one launch each) on a 2e8-record C4 stream
mkdir -p gpurun_out
B="python bench.py --reads 100000 --cpu-sample 0 --no-e2e --cli-reads 0 --cov-cpu-sample 0 --steps 1 --warmup 1 --cov-records 200000000"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-cov_accumulate_kernel|cov_bundle_kernel}" -c ${KCOUNT:-2} -f -o gpurun_out/${TAG:-r2z}_cov_prof $B > gpurun_out/${TAG:-r2z}_cov_prof.log 2>&1
ls -la gpurun_out | tail -4
