"""GPU: the sharded tiecov / tiebrush paths of tiebrush_b200/shard.py with the CUDA library as the per-shard engine
(all shards run one after the other on cuda:0; the collectives are covered by tests/test_shard.py with gloo and by
tools/shard_nccl_check.py under torchrun on >1 GPU)."""
import numpy as np
import pytest

import test_shard as T
from oracle import oracle
from tiebrush_b200 import api, shard, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with api.Context(device=0, n_samples=8) as c:
        yield c


@pytest.mark.parametrize("world", [2, 8])
def test_gpu_coverage_shards(ctx, world, monkeypatch):
    cols = T._cov_stream(40000, 21, chroms=3, n_tx=90)
    monkeypatch.setattr(T, "_cov_compute", lambda c, wr, wj: ctx.coverage_window(c, want_runs=wr, want_juncs=wj))
    cuts = shard.cov_cuts(cols, world)
    cuts[0] = (cuts[0][0], cuts[0][1] + 53)
    T._assert_cov(T._simulate_cov(cols, cuts), oracle.coverage(cols))


@pytest.mark.parametrize("world", [2, 8])
def test_gpu_collapse_shards(ctx, world):
    cols, run_off, _ = synth.cohort_window(8, 6000, seed=13, n_tx=40, device="cpu")
    host = synth.to_host(cols)
    exp = oracle.collapse(host, run_off)
    cuts = shard.collapse_cuts(host, run_off, world)
    bounds = [None] + cuts + [None]
    parts = [shard.collapse_shard_local(lambda c, ro: ctx.collapse_window(c, ro), host, run_off, bounds[g], bounds[g + 1]) for g in range(world)]
    for key in ("rep_index", "yc", "yx", "yd"):
        got = np.concatenate([np.asarray(p[key]) for p in parts])
        assert np.array_equal(got.astype(np.float64), np.asarray(exp[key]).astype(np.float64)), key


@pytest.mark.parametrize("world", [2, 8])
def test_gpu_sample_heatmap_shards(ctx, world):
    """Sharded tiecov -s with the device path as the per-shard engine: cuts inside bundles, rows stitched at the cuts."""
    cols = T._sample_cols(30000, 23, chroms=2, n_tx=60)
    exp = oracle.sample_heatmap(cols)
    cuts = [(t, p + 29) for t, p in shard.cov_cuts(cols, world)]
    bounds = [None] + cuts + [None]
    parts = [shard.sample_shard_local(ctx.sample_window, cols, bounds[g], bounds[g + 1]) for g in range(world)]
    for a, b in zip(shard._stitch_sample(parts), exp):
        assert np.array_equal(a, b)


def test_shard_nccl_c4_multi_gpu():
    """The C++ / NCCL sharded tiecov path (tc_shard_coverage + tc_shard_gather) on every GPU of the box: gathered rows equal
    the single-GPU stream and the oracle (tools/shard_nccl_c4_check.py under torchrun). Needs >= 2 GPUs."""
    import os, subprocess, sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(ngpu, 8)}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(root, "tools", "shard_nccl_c4_check.py"), "300000"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and "SHARD_NCCL_C4_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
