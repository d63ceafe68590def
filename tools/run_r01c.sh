(time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r01c_bench_10M.json 2> gpurun_out/r01c_bench_10M.err; tail -5 gpurun_out/r01c_bench_10M.err
timeout 600 python -m pytest tests/test_gpu_shard.py -x -q 2>&1 | tail -5
