#!/bin/bash
mkdir -p gpurun_out
(time timeout 2400 python -m pytest tests -q -m gpu -x) > gpurun_out/${TAG:-r2s}_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/${TAG:-r2s}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
TAG=${TAG:-r2s} bash tools/gpu_bench_default.sh
