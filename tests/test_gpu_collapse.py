"""GPU parity tests (through the C ABI): merge + collapse kernels vs golden outputs of the compiled reference
and vs the oracle on seeded synthetic cohorts."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def gpu_collapse(cols, run_off, tid=0, file_merged=None, **opts):
    from tiebrush_b200 import api
    k = len(run_off) - 1
    with api.Context(device=0, n_samples=k, **opts) as ctx:
        return ctx.collapse_window(cols, run_off, tid=tid, file_merged=file_merged)


@pytest.mark.parametrize("case", H.case_names("collapse_random.npz"))
def test_collapse_random(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_random.npz", case)
    H.assert_collapse_equal(H.run_collapse(gpu_collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("case", H.case_names("collapse_fixture.npz"))
def test_collapse_fixture(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_fixture.npz", case)
    H.assert_collapse_equal(H.run_collapse(gpu_collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("case", H.case_names("collapse_merged.npz"))
def test_collapse_merged(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_merged.npz", case)
    H.assert_collapse_equal(H.run_collapse(gpu_collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("n_tx,k,reads", [(40, 12, 20000), (2000, 40, 5000), (3, 5, 30000)])
def test_collapse_synthetic_vs_oracle(mode, n_tx, k, reads):
    """Seeded synthetic cohorts (deep pile-ups when n_tx is tiny) against the oracle, bit for bit."""
    from oracle import oracle
    from tiebrush_b200 import synth
    cols, run_off, pr = synth.cohort_window(k, reads, seed=3, n_tx=n_tx, device="cpu")
    host = synth.to_host(cols)
    got = gpu_collapse(host, run_off, mode=mode)
    exp = oracle.collapse(host, run_off, mode=mode)
    assert got["n_kept"] == exp["n_kept"]
    for key in ("rep_index", "yc", "yx", "yd"):
        assert np.array_equal(np.asarray(got[key]), exp[key]), key


def test_collapse_device_resident_properties():
    """BASELINE-shaped cohort kept in HBM: size-independent properties (sum YC == records kept, reps are
    members with the right position order, YX <= min(k, YC)) plus oracle equality on the host copy."""
    import torch
    from oracle import oracle
    from tiebrush_b200 import api, synth
    k, reads = 50, 40000
    cols, run_off, pr = synth.cohort_window(k, reads, seed=5, n_tx=500, device="cuda")
    with api.Context(device=0, n_samples=k) as ctx:
        got = ctx.collapse_window(cols, run_off, pos_range=pr)
    rep = got["rep_index"].cpu().numpy().view(np.uint32); yc = got["yc"].cpu().numpy(); yx = got["yx"].cpu().numpy().view(np.uint32)
    assert got["n_kept"] == k * reads and float(yc.sum()) == k * reads
    pos = cols["pos"].cpu().numpy()
    assert (np.diff(pos[rep]) >= 0).all()
    assert (yx <= np.minimum(k, yc)).all() and (yx >= 1).all()
    host = synth.to_host(cols)
    exp = oracle.collapse(host, run_off)
    assert np.array_equal(rep, exp["rep_index"]) and np.array_equal(yc, exp["yc"]) and np.array_equal(yx, exp["yx"])
    assert np.array_equal(got["yd"].cpu().numpy(), exp["yd"])
