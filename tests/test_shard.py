"""Coordinate sharding: host-side edge logic of tiebrush_b200/shard.py on CPU.

The per-record compute is injected: here it is the oracle (tests may use it as the checker's engine), on the GPU box it
is api.Context (tests/test_gpu_shard.py). The multi-rank paths run with the gloo backend, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest

from oracle import oracle
from tiebrush_b200 import shard, synth


def _cov_compute(cols, want_runs, want_juncs):
    return oracle.coverage(cols, want_runs=want_runs, want_juncs=want_juncs)


def _col_compute(cols, run_off):
    return oracle.collapse(cols, run_off)


def _cov_stream(n, seed, chroms=1, n_tx=40):
    return synth.to_host(synth.coverage_stream(n, seed=seed, n_tx=n_tx, chroms=chroms))


def _simulate_cov(cols, cuts):
    """All ranks in one process: what coverage_sharded does, without the collectives."""
    world = len(cuts) + 1
    bounds = [None] + list(cuts) + [None]
    owns = [shard.cov_slice(cols, bounds[g], bounds[g + 1]) for g in range(world)]
    fars = []
    for g in range(world):
        hi = bounds[g + 1]
        e = shard.ref_end(owns[g])
        far = np.nonzero((owns[g]["tid"] == hi[0]) & (e > hi[1]))[0] if hi is not None else np.zeros(0, np.int64)
        fars.append(shard._unpack_cov(shard._pack_cov(shard._cov_take(owns[g], far))))
    parts = []
    for g in range(world):
        lo = bounds[g]
        halo = []
        if lo is not None:
            for h in range(g):
                c = fars[h]
                if len(c["pos"]):
                    halo.append(shard._cov_take(c, np.nonzero((c["tid"] == lo[0]) & (shard.ref_end(c) > lo[1]))[0]))
        halo_in = shard._cov_concat(halo)
        if len(halo_in["pos"]) > 1:
            halo_in = shard._cov_take(halo_in, np.argsort(halo_in["pos"], kind="stable"))
        parts.append(shard.coverage_shard_local(_cov_compute, owns[g], lo, bounds[g + 1], halo_in))
    return shard._assemble(parts, cuts)


def _assert_cov(got, exp):
    for a, b, name in zip(got[0], exp["runs"], ("tid", "start0", "end0", "value")):
        assert np.array_equal(np.asarray(a), b), f"runs.{name}"
    for a, b, name in zip(got[1], exp["juncs"], ("tid", "start", "end", "strand", "value")):
        assert np.array_equal(np.asarray(a), b), f"juncs.{name}"


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_coverage_shards_balanced_cuts(seed, world):
    cols = _cov_stream(6000, seed, chroms=2 if seed % 2 else 1)
    exp = oracle.coverage(cols)
    _assert_cov(_simulate_cov(cols, shard.cov_cuts(cols, world)), exp)


@pytest.mark.parametrize("seed", range(6))
def test_coverage_shards_arbitrary_cuts(seed):
    """Cuts at random coordinates: inside reads, inside introns, between bundles, before the first / after the last record,
    and several cuts inside one long bundle (a rank that owns nothing but sees the halo)."""
    rng = np.random.default_rng(seed)
    cols = _cov_stream(3000, 100 + seed, n_tx=12)
    exp = oracle.coverage(cols)
    lo, hi = int(cols["pos"].min()), int(cols["pos"].max())
    for _ in range(6):
        npos = rng.integers(1, 6)
        pick = np.sort(rng.choice(cols["pos"], npos) + rng.integers(-40, 200, npos))
        cuts = [(0, int(np.clip(p, lo - 5, hi + 500))) for p in pick]
        _assert_cov(_simulate_cov(cols, cuts), exp)


def test_coverage_shards_dense_cuts_inside_one_bundle():
    cols = _cov_stream(400, 7, n_tx=2)
    exp = oracle.coverage(cols)
    p0 = int(np.median(cols["pos"]))
    cuts = [(0, p0 + 7 * i) for i in range(10)]
    _assert_cov(_simulate_cov(cols, cuts), exp)


def test_gap_at_or_before_is_a_gap():
    cols, run_off, _ = synth.cohort_window(5, 800, seed=3, n_tx=20, device="cpu")
    host = synth.to_host(cols)
    e = shard.ref_end(host)
    for cut in np.quantile(host["pos"], [0.1, 0.35, 0.5, 0.8]).astype(int):
        g = shard.gap_at_or_before(host, run_off, int(cut))
        assert g <= cut
        m = host["pos"] < g
        assert not m.any() or e[m].max() <= g


@pytest.mark.parametrize("world", [2, 4])
def test_collapse_shards_equal_whole_window(world):
    cols, run_off, _ = synth.cohort_window(6, 1500, seed=5, n_tx=25, device="cpu")
    host = synth.to_host(cols)
    exp = oracle.collapse(host, run_off)
    cuts = shard.collapse_cuts(host, run_off, world)
    bounds = [None] + cuts + [None]
    parts = [shard.collapse_shard_local(_col_compute, host, run_off, bounds[g], bounds[g + 1]) for g in range(world)]
    for key in ("rep_index", "yc", "yx", "yd"):
        got = np.concatenate([p[key] for p in parts])
        assert np.array_equal(got.astype(np.int64) if key == "rep_index" else got, exp[key].astype(np.int64) if key == "rep_index" else exp[key]), key


# ---- torch.distributed, gloo, world_size > 1 ----------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, what, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if what == "cov":
            cols = _cov_stream(5000, 11, chroms=2)
            cuts = shard.cov_cuts(cols, world)
            cuts[0] = (cuts[0][0], cuts[0][1] + 37)          # off the balanced point: inside reads
            bounds = [None] + cuts + [None]
            own = shard.cov_slice(cols, bounds[rank], bounds[rank + 1])
            runs, juncs = shard.coverage_sharded(_cov_compute, own, bounds[rank], bounds[rank + 1], cuts)
            if rank == 0:
                exp = oracle.coverage(cols)
                ok = all(np.array_equal(a, b) for a, b in zip(runs, exp["runs"])) and all(np.array_equal(a, b) for a, b in zip(juncs, exp["juncs"]))
                q.put(("cov", bool(ok), len(runs[0]), len(juncs[0])))
        elif what == "sample":
            cols = _cov_stream(6000, 13, chroms=2)
            cols["yx"] = np.random.default_rng(13).integers(1, 40, size=6000).astype(np.int32)
            cuts = [(t, p + 41) for t, p in shard.cov_cuts(cols, world)]   # inside bundles
            got = shard.sample_sharded(oracle.sample_heatmap, cols, cuts)
            if rank == 0:
                exp = oracle.sample_heatmap(cols)
                q.put(("sample", bool(all(np.array_equal(a, b) for a, b in zip(got, exp))), len(got[0]), 0))
        else:
            cols, run_off, _ = synth.cohort_window(5, 2000, seed=9, n_tx=25, device="cpu")
            host = synth.to_host(cols)
            cuts = shard.collapse_cuts(host, run_off, world)
            got = shard.collapse_sharded(_col_compute, host, run_off, cuts)
            if rank == 0:
                exp = oracle.collapse(host, run_off)
                ok = all(np.array_equal(np.asarray(got[k]).astype(np.int64), np.asarray(exp[k]).astype(np.int64)) for k in ("rep_index", "yx", "yd")) \
                    and np.array_equal(got["yc"], exp["yc"])
                q.put(("col", bool(ok), len(got["rep_index"]), 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("what,world", [("cov", 2), ("cov", 3), ("col", 2), ("sample", 2), ("sample", 3)])
def test_sharded_gloo(what, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, what, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    tag, ok, a, b = q.get(timeout=10)
    assert tag == what and ok and a > 0


def test_compact_cigar_wire_format_round_trip():
    """api.compact_cigar_columns: n_cigar8 / cigar16 / cigar_ext expand back to the wide CSR columns (the rule the device applies)."""
    import numpy as np
    from tiebrush_b200 import api, synth
    cols, run_off, _ = synth.cohort_window(6, 5000, seed=2, n_tx=200, device="cpu")
    host = synth.to_host(cols)
    n8, c16, ext = api.compact_cigar_columns(host["cig_off"], host["cigar"])
    off = np.zeros(len(n8) + 1, np.uint32); off[1:] = np.cumsum(n8)
    esc = (c16 >> 4) == 0xFFF
    ln = (c16 >> 4).astype(np.uint32)
    ln[esc] = ext
    wide = (c16 & 0xF).astype(np.uint32) | (ln << 4)
    assert np.array_equal(off, host["cig_off"]) and np.array_equal(wide, host["cigar"]) and esc.sum() == len(ext) > 0
    t8, t16, text = api.compact_cigar_columns(cols["cig_off"], cols["cigar"])   # torch path
    assert np.array_equal(t8.numpy(), n8) and np.array_equal(t16.numpy().view(np.uint16), c16) and np.array_equal(text.numpy().view(np.uint32), ext)


def test_prefix_slice_is_a_coordinate_prefix_of_every_file():
    """synth.prefix_slice (the bench's CPU-baseline sample and the full-size tests' oracle prefix): every file contributes
    exactly its records below one coordinate, CSR columns stay consistent, and the oracle on the slice equals the head of
    the oracle on the whole window (groups and YD only look left)."""
    import numpy as np
    from oracle import oracle
    from tiebrush_b200 import synth
    cols, run_off, _ = synth.cohort_window(7, 3000, seed=8, n_tx=60, device="cpu")
    host = synth.to_host(cols)
    sub, sub_off = synth.prefix_slice(cols, run_off, 6000)
    hi = int(sub["pos"].max()) + 1
    assert len(sub_off) == len(run_off) and sub_off[-1] == len(sub["pos"]) == int(sub["cig_off"].shape[0]) - 1
    for f in range(len(run_off) - 1):
        a, b = int(run_off[f]), int(run_off[f + 1])
        want = host["pos"][a:b]
        m = int(sub_off[f + 1] - sub_off[f])
        assert np.array_equal(sub["pos"][sub_off[f]:sub_off[f + 1]], want[:m]) and (m == b - a or want[m] >= hi)
    full, part = oracle.collapse(host, run_off), oracle.collapse(sub, sub_off)
    g = len(part["rep_index"])
    assert 0 < g < len(full["rep_index"])
    for key in ("yc", "yx", "yd"):
        assert np.array_equal(part[key], full[key][:g]), key


def _sample_cols(n, seed, chroms=1, n_tx=40):
    cols = _cov_stream(n, seed, chroms=chroms, n_tx=n_tx)
    cols["yx"] = np.random.default_rng(seed).integers(1, 40, size=n).astype(np.int32)
    return cols


@pytest.mark.parametrize("seed,world,chroms", [(1, 2, 1), (2, 3, 2), (3, 8, 3)])
def test_sample_heatmap_shards_equal_whole_stream(seed, world, chroms):
    """Sharded tiecov -s (all ranks simulated in one process): balanced cuts and cuts nudged into the middle of bundles give
    the rows of the unsharded stream."""
    cols = _sample_cols(20000, seed, chroms=chroms)
    exp = oracle.sample_heatmap(cols)
    for shift in (0, 37):
        cuts = [(t, p + shift) for t, p in shard.cov_cuts(cols, world)]
        bounds = [None] + cuts + [None]
        parts = [shard.sample_shard_local(oracle.sample_heatmap, cols, bounds[g], bounds[g + 1]) for g in range(world)]
        got = shard._stitch_sample(parts)
        for a, b in zip(got, exp):
            assert np.array_equal(a, b), (seed, world, shift)


def _tiny_stream(recs):
    """recs = [(pos, M length)] on tid 0, YX = 3, 7, 11 ..."""
    n = len(recs)
    return dict(tid=np.zeros(n, np.int32), pos=np.asarray([p for p, _ in recs], np.int32), yc=np.ones(n, np.float32),
                strand=np.full(n, ord("."), np.uint8), cig_off=np.arange(n + 1, dtype=np.uint32),
                cigar=np.asarray([(l << 4) | 0 for _, l in recs], np.uint32), yx=np.asarray([3 + 4 * i for i in range(n)], np.int32))


@pytest.mark.parametrize("recs,cuts", [
    ([(0, 100), (200, 50)], [(0, 50)]),                       # the bundle reaches past the cut, no record of it starts there
    ([(0, 100), (200, 50)], [(0, 50), (0, 80)]),              # ... and a rank with no record of its own in between
    ([(0, 100), (200, 50)], [(0, 150)]),                      # cut in the gap
    ([(0, 300), (10, 20), (400, 30)], [(0, 5), (0, 100), (0, 250)]),
    ([(0, 300), (10, 20), (400, 30)], [(0, 11), (0, 350)]),
    ([(0, 300), (10, 20), (400, 30)], [(0, 299), (0, 300), (0, 301)]),
])
def test_sample_heatmap_shards_sparse_streams_and_empty_ranks(recs, cuts):
    """a bundle open at a cut with no further record after it, ranks without records, cuts in gaps."""
    cols = _tiny_stream(recs)
    exp = oracle.sample_heatmap(cols)
    bounds = [None] + cuts + [None]
    parts = [shard.sample_shard_local(oracle.sample_heatmap, cols, bounds[g], bounds[g + 1]) for g in range(len(cuts) + 1)]
    got = shard._stitch_sample(parts)
    for a, b in zip(got, exp):
        assert np.array_equal(a, b), (recs, cuts, got, exp)
