#!/usr/bin/env bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/last_pytest.log 2>&1; tail -4 gpurun_out/last_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
