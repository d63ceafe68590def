#!/usr/bin/env python
"""Multi-GPU check of tiebrush_b200/shard.py over NCCL (run under torchrun on the GPU box, e.g.
`python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_nccl_check.py`):
sharded tiecov with the halo exchange + ordered gather and sharded tiebrush, each rank on its own GPU, compared on rank 0
with the single-window CUDA result and the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tiebrush_b200 import api, shard, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(device=local, n_samples=16)
cols = synth.to_host(synth.coverage_stream(400000, seed=3, n_tx=400, chroms=3))
cuts = shard.cov_cuts(cols, world)
cuts = [(t, p + 61) for t, p in cuts]
bounds = [None] + cuts + [None]
own = shard.cov_slice(cols, bounds[rank], bounds[rank + 1])
runs, juncs = shard.coverage_sharded(lambda c, wr, wj: ctx.coverage_window(c, want_runs=wr, want_juncs=wj), own, bounds[rank], bounds[rank + 1], cuts)
ok = True
if rank == 0:
    from oracle import oracle
    exp = oracle.coverage(cols)
    ok &= all(np.array_equal(a, b) for a, b in zip(runs, exp["runs"])) and all(np.array_equal(a, b) for a, b in zip(juncs, exp["juncs"]))
    print(f"sharded tiecov over {world} GPUs: {len(runs[0])} runs, {len(juncs[0])} junctions, equal to the oracle: {ok}")
c2, run_off, _ = synth.cohort_window(16, 50000, seed=4, n_tx=300, device="cpu")
host = synth.to_host(c2)
got = shard.collapse_sharded(lambda c, ro: ctx.collapse_window(c, ro), host, run_off, shard.collapse_cuts(host, run_off, world))
if rank == 0:
    exp = oracle.collapse(host, run_off)
    ok2 = all(np.array_equal(np.asarray(got[k]).astype(np.float64), np.asarray(exp[k]).astype(np.float64)) for k in ("rep_index", "yc", "yx", "yd"))
    print(f"sharded tiebrush over {world} GPUs: {len(got['rep_index'])} groups, equal to the oracle: {ok2}")
    ok &= ok2
ctx.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
