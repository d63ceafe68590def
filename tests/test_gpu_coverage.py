"""GPU parity tests (through the C ABI): tiecov coverage / junction / bedgraph kernels vs the oracle and
vs golden outputs of the compiled reference."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from tiebrush_b200 import api
    with api.Context(device=0) as c:
        yield c


@pytest.mark.parametrize("case", H.coverage_case_names())
def test_coverage_golden(ctx, case):
    cols, runs, juncs = H.load_coverage_case(case)
    got = ctx.coverage_window(cols)
    H.assert_coverage_equal(got, runs, juncs, case)


@pytest.mark.parametrize("case", H.coverage_case_names()[:4])
def test_coverage_matches_oracle_exactly(ctx, case):
    from oracle import oracle
    cols, _, _ = H.load_coverage_case(case)
    got, exp = ctx.coverage_window(cols), oracle.coverage(cols)
    for a, b in zip(got["runs"], exp["runs"]):
        assert np.array_equal(a, b)
    for a, b in zip(got["juncs"], exp["juncs"]):
        assert np.array_equal(a, b)


def test_coverage_only_runs_or_only_juncs(ctx):
    cols, runs, juncs = H.load_coverage_case(H.coverage_case_names()[0])
    r = ctx.coverage_window(cols, want_juncs=False)
    assert r["n_runs"] == len(runs[0]) and r["n_juncs"] == 0
    j = ctx.coverage_window(cols, want_runs=False)
    assert j["n_juncs"] == len(juncs[0]) and j["n_runs"] == 0


def test_coverage_rejects_unsupported_ops(ctx):
    cols = dict(tid=np.zeros(2, np.int32), pos=np.asarray([10, 12], np.int32), yc=np.ones(2, np.float32),
                strand=np.asarray([ord(".")] * 2, np.uint8), cig_off=np.asarray([0, 1, 2], np.uint32),
                cigar=np.asarray([(5 << 4) | 0, (5 << 4) | 7], np.uint32))
    with pytest.raises(ValueError):
        ctx.coverage_window(cols)
    # junction-only runs never look at the ops (addCov is not called, tiecov.cpp:486-488)
    assert ctx.coverage_window(cols, want_runs=False)["n_juncs"] == 0


def test_coverage_empty(ctx):
    cols = dict(tid=np.zeros(0, np.int32), pos=np.zeros(0, np.int32), yc=np.zeros(0, np.float32), strand=np.zeros(0, np.uint8),
                cig_off=np.zeros(1, np.uint32), cigar=np.zeros(0, np.uint32))
    r = ctx.coverage_window(cols)
    assert r["n_runs"] == 0 and r["n_juncs"] == 0


def test_coverage_device_resident_large_random(ctx):
    """Size-independent properties on a large synthetic stream kept in HBM: total covered weight is conserved,
    runs are sorted, disjoint, never zero, and adjacent equal-valued runs only meet at bundle edges."""
    import torch
    from tiebrush_b200 import synth
    cols = synth.coverage_stream(n=2_000_000, seed=7, device="cuda")
    out = ctx.coverage_window(cols)
    t, s, e, v = (x.cpu().numpy() for x in out["runs"])
    assert len(t) > 0 and (v != 0).all() and (e > s).all()
    same = t[1:] == t[:-1]
    assert (s[1:][same] >= e[:-1][same]).all()
    total = float(((e - s).astype(np.float64) * v).sum())
    assert total == synth.covered_weight(cols)
    # against the oracle on a prefix that the CPU finishes quickly (cut at a bundle boundary)
    from oracle import oracle
    host = synth.to_host(cols)
    cut = synth.bundle_cut(host, 200_000)
    sub = synth.take_prefix(host, cut)
    exp = oracle.coverage(sub)
    got = ctx.coverage_window(sub)
    for a, b in zip(got["runs"], exp["runs"]):
        assert np.array_equal(np.asarray(a), b)
    for a, b in zip(got["juncs"], exp["juncs"]):
        assert np.array_equal(np.asarray(a), b)
