nvidia-smi --query-gpu=name,memory.total --format=csv; free -g | head -2; nproc
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r01b_pytest.log 2>&1; tail -3 gpurun_out/r01b_pytest.log
timeout 600 python bench.py --reads 2000000 --cov-records 50000000 --steps 3 --warmup 3 > gpurun_out/r01b_bench_2M.json 2> gpurun_out/r01b_bench_2M.err; tail -2 gpurun_out/r01b_bench_2M.err
MEM=$(free -g | awk '/Mem:/{print $7}')
if [ "$MEM" -gt 150 ]; then (time timeout 900 python bench.py --steps 3 --warmup 3) > gpurun_out/r01b_bench_10M.json 2> gpurun_out/r01b_bench_10M.err; tail -5 gpurun_out/r01b_bench_10M.err; fi
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_ref.json 2>&1
KCOUNT=12 KREGEX='col_|yd_|cov_|junc_|ord_' bash tools/prof.sh r01b > gpurun_out/r01b_prof_sh.log 2>&1
