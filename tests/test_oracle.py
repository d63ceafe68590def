"""CPU tests: the C restatement (oracle/tb_oracle.c) against golden vectors produced by the UNMODIFIED
compiled reference (tests/golden/make_golden.py). This is what pins the oracle."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from oracle import oracle
from tiebrush_b200 import sam


@pytest.mark.parametrize("case", H.case_names("collapse_random.npz"))
def test_collapse_random(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_random.npz", case)
    H.assert_collapse_equal(H.run_collapse(oracle.collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("case", H.case_names("collapse_fixture.npz"))
def test_collapse_fixture(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_fixture.npz", case)
    H.assert_collapse_equal(H.run_collapse(oracle.collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("case", H.case_names("collapse_merged.npz"))
def test_collapse_merged(case):
    files, opts, fm, exp = H.load_collapse_case("collapse_merged.npz", case)
    H.assert_collapse_equal(H.run_collapse(oracle.collapse, files, opts, fm), exp, case)


@pytest.mark.parametrize("case", H.coverage_case_names())
def test_coverage(case):
    cols, runs, juncs = H.load_coverage_case(case)
    H.assert_coverage_equal(oracle.coverage(cols), runs, juncs, case)


def test_coverage_rejects_unsupported_ops():
    cols = dict(tid=np.zeros(1, np.int32), pos=np.asarray([10], np.int32), yc=np.ones(1, np.float32),
                strand=np.asarray([ord(".")], np.uint8), cig_off=np.asarray([0, 1], np.uint32),
                cigar=np.asarray([(5 << 4) | 7], np.uint32))
    with pytest.raises(ValueError):
        oracle.coverage(cols)


REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not (os.path.isdir(REF_TEST) and os.path.exists(os.path.join(oracle.REF_DIR, "tiebrush"))),
                    reason="needs /root/reference and oracle/_ref (build container only)")
def test_full_fixtures_against_live_reference(tmp_path):
    """All 20 sample BAMs of the reference's own tests: oracle vs the compiled reference run live."""
    paths = [f"{REF_TEST}/t1/t1s{i}.bam" for i in range(10)] + [f"{REF_TEST}/t2/t2s{i}.bam" for i in range(10)]
    out = str(tmp_path / "o.bam")
    r = subprocess.run([os.path.join(oracle.REF_DIR, "tiebrush"), "-o", out] + paths, check=True, capture_output=True, text=True)
    assert r.stderr.startswith("659832 input records written as 9491")
    hts = os.path.join(oracle.REF_DIR, "htsfile")
    exp_txt = subprocess.run([hts, "-c", out], check=True, capture_output=True, text=True).stdout
    expR, ids = sam.parse_sam(exp_txt)
    expc = sam.to_columns(expR)
    files = [sam.to_columns(sam.parse_sam(subprocess.run([hts, "-c", p], check=True, capture_output=True, text=True).stdout, ids)[0])
             for p in paths]
    got = H.run_collapse(oracle.collapse, files, {})
    H.assert_collapse_equal(got, dict(tid=expc["tid"], lhash=expc["lhash"], yc=expc["yc_in"], yx=expc["yx_in"],
                                      yd=expc["yd_in"], n_kept=659832), "t1+t2")
    # tiecov on the collapsed output
    subprocess.run([os.path.join(oracle.REF_DIR, "tiecov"), "-c", str(tmp_path / "k.cov"), "-j", str(tmp_path / "k.j"), out], check=True)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as mg
    runs = mg.parse_bedgraph(open(tmp_path / "k.cov.bedgraph").read(), ids)
    juncs = mg.parse_bed(open(tmp_path / "k.j.bed").read(), ids)
    assert len(runs[0]) == 9043 and len(juncs[0]) == 18
    cols = dict(expc)
    cols["yc"] = np.where(expc["has_yc"], expc["yc_in"], np.float32(1)).astype(np.float32)
    H.assert_coverage_equal(oracle.coverage(cols), runs, juncs, "t1+t2 tiecov")
