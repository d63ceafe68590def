#!/bin/bash
# closing measurements on one GPU: full GPU suite, smoke, the driver's default line, the reference arm
mkdir -p gpurun_out
TAG=${TAG:-r2z} bash tools/gpu_full_suite.sh
(time timeout 900 python bench.py --impl reference) > gpurun_out/${TAG:-r2z}_reference_arm.json 2> gpurun_out/${TAG:-r2z}_reference_arm.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${TAG:-r2z}_reference_arm.json
