"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/libtb_oracle.so (tb_oracle.c).

Takes the same column dicts the product's Python mirror takes (tiebrush_b200.sam.to_columns layout) so
that parity tests feed identical inputs to both sides."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class _In(C.Structure):
    _fields_ = [("n", C.c_int64), ("n_files", C.c_int32), ("tid", C.c_int32), ("run_off", C.c_void_p),
                ("file_merged", C.c_void_p), ("pos", C.c_void_p), ("flag", C.c_void_p), ("mapq", C.c_void_p),
                ("strand", C.c_void_p), ("nh", C.c_void_p), ("cig_off", C.c_void_p), ("cigar", C.c_void_p),
                ("md_off", C.c_void_p), ("md", C.c_void_p), ("qhash", C.c_void_p), ("yc_in", C.c_void_p),
                ("yx_in", C.c_void_p), ("yd_in", C.c_void_p), ("on_device", C.c_int32),
                ("n_cig", C.c_int64), ("n_md", C.c_int64), ("pos_lo", C.c_int32), ("pos_hi", C.c_int32)]


class _Out(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("n_groups", C.c_int64), ("n_kept", C.c_int64), ("rep_index", C.c_void_p),
                ("yc", C.c_void_p), ("yx", C.c_void_p), ("yd", C.c_void_p), ("on_device", C.c_int32)]


class _Opts(C.Structure):
    _fields_ = [("mode", C.c_int), ("flag_mask", C.c_uint32), ("max_nh", C.c_int), ("min_qual", C.c_int),
                ("keep_bits", C.c_int), ("collapse_same", C.c_int)]


class _CovIn(C.Structure):
    _fields_ = [("n", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("yc", C.c_void_p), ("strand", C.c_void_p),
                ("cig_off", C.c_void_p), ("cigar", C.c_void_p), ("on_device", C.c_int32), ("n_cig", C.c_int64)]


class _Runs(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("n_runs", C.c_int64), ("tid", C.c_void_p), ("start0", C.c_void_p),
                ("end0", C.c_void_p), ("value", C.c_void_p), ("on_device", C.c_int32)]


class _Juncs(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("n_juncs", C.c_int64), ("tid", C.c_void_p), ("start", C.c_void_p),
                ("end", C.c_void_p), ("strand", C.c_void_p), ("value", C.c_void_p), ("on_device", C.c_int32)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libtb_oracle.so"])
    subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libtb_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(HERE, "tb_oracle.c")):
            subprocess.check_call(["make", "-s", "-C", HERE, "libtb_oracle.so"])
        _lib = C.CDLL(path)
        _lib.tbo_collapse.restype = C.c_int
        _lib.tbo_coverage.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def collapse(cols: dict, run_off, tid=0, mode=0, flag_mask=0, max_nh=0x7FFFFFFF, min_qual=-1, keep_bits=0,
             collapse_same=0, file_merged=None):
    """Oracle collapse of one window. Returns dict(rep_index, yc, yx, yd, n_kept)."""
    n = len(cols["pos"])
    run_off = _c(run_off, np.int64)
    k = len(run_off) - 1
    a = dict(pos=_c(cols["pos"], np.int32), flag=_c(cols["flag"], np.uint16), mapq=_c(cols["mapq"], np.uint8),
             strand=_c(cols["strand"], np.uint8), nh=_c(cols["nh"], np.uint16), cig_off=_c(cols["cig_off"], np.uint32),
             cigar=_c(cols["cigar"], np.uint32))
    md_off = _c(cols["md_off"], np.uint32) if "md_off" in cols else np.zeros(n + 1, np.uint32)
    md = _c(cols["md"], np.uint8) if "md" in cols else np.zeros(0, np.uint8)
    qh = _c(cols["qhash"], np.uint64) if "qhash" in cols else None
    fm = _c(file_merged, np.uint8) if file_merged is not None else None
    yc_in = _c(cols["yc_in"], np.float32) if (fm is not None and "yc_in" in cols) else None
    yx_in = _c(cols["yx_in"], np.int32) if (fm is not None and "yx_in" in cols) else None
    yd_in = _c(cols["yd_in"], np.int32) if (fm is not None and "yd_in" in cols) else None
    sin = _In(n, k, tid, _p(run_off), _p(fm), _p(a["pos"]), _p(a["flag"]), _p(a["mapq"]), _p(a["strand"]), _p(a["nh"]),
              _p(a["cig_off"]), _p(a["cigar"]), _p(md_off), _p(md), _p(qh), _p(yc_in), _p(yx_in), _p(yd_in), 0,
              int(a["cig_off"][-1]) if n else 0, int(md_off[-1]) if n else 0, 0, 0)
    cap = max(n, 1)
    rep = np.zeros(cap, np.uint32); yc = np.zeros(cap, np.float32); yx = np.zeros(cap, np.uint32); yd = np.zeros(cap, np.int32)
    out = _Out(cap, 0, 0, _p(rep), _p(yc), _p(yx), _p(yd), 0)
    o = _Opts(mode, flag_mask, max_nh, min_qual, keep_bits, collapse_same)
    rc = lib().tbo_collapse(C.byref(sin), C.byref(o), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"tbo_collapse rc={rc}")
    g = out.n_groups
    return dict(rep_index=rep[:g].copy(), yc=yc[:g].copy(), yx=yx[:g].copy(), yd=yd[:g].copy(), n_kept=int(out.n_kept))


def coverage(cols: dict, want_runs=True, want_juncs=True, cap_runs=None):
    """Oracle tiecov -c/-j of one window. Returns dict(runs=(tid,start0,end0,value), juncs=(tid,start,end,strand,value))."""
    n = len(cols["pos"])
    a = dict(tid=_c(cols["tid"], np.int32), pos=_c(cols["pos"], np.int32), yc=_c(cols["yc"], np.float32),
             strand=_c(cols["strand"], np.uint8), cig_off=_c(cols["cig_off"], np.uint32), cigar=_c(cols["cigar"], np.uint32))
    cin = _CovIn(n, _p(a["tid"]), _p(a["pos"]), _p(a["yc"]), _p(a["strand"]), _p(a["cig_off"]), _p(a["cigar"]), 0, int(a["cig_off"][-1]) if n else 0)
    ncig = int(a["cig_off"][-1]) if n else 0
    capr = cap_runs if cap_runs is not None else 2 * ncig + 16
    capj = ncig + 16
    rt = np.zeros(capr, np.int32); rs = np.zeros(capr, np.int32); re = np.zeros(capr, np.int32); rv = np.zeros(capr, np.float64)
    jt = np.zeros(capj, np.int32); js = np.zeros(capj, np.int32); je = np.zeros(capj, np.int32)
    jc = np.zeros(capj, np.uint8); jv = np.zeros(capj, np.float64)
    runs = _Runs(capr, 0, _p(rt), _p(rs), _p(re), _p(rv), 0)
    juncs = _Juncs(capj, 0, _p(jt), _p(js), _p(je), _p(jc), _p(jv), 0)
    bad = C.c_int64(-1)
    rc = lib().tbo_coverage(C.byref(cin), C.byref(runs) if want_runs else None, C.byref(juncs) if want_juncs else None,
                            C.byref(bad))
    if rc == 2:
        raise ValueError(f"unsupported CIGAR op in record {bad.value}")
    if rc != 0:
        raise RuntimeError(f"tbo_coverage rc={rc}")
    r, j = runs.n_runs, juncs.n_juncs
    return dict(runs=(rt[:r].copy(), rs[:r].copy(), re[:r].copy(), rv[:r].copy()),
                juncs=(jt[:j].copy(), js[:j].copy(), je[:j].copy(), jc[:j].copy(), jv[:j].copy()))


def sample_heatmap(cols: dict, cap_runs=None):
    """Oracle tiecov -s of one stream (GROUNDWORK for SURVEY §8f.2, no device path yet): rows (tid, start0, end0, ival) of the
    sample-count heat-map; `cols` needs tid, pos, yx (YX tag, 1 when absent), cig_off, cigar. hval of a row is
    heatmap_value(ival, n_samples)."""
    n = len(cols["pos"])
    a = dict(tid=_c(cols["tid"], np.int32), pos=_c(cols["pos"], np.int32), yx=_c(cols["yx"], np.int32),
             cig_off=_c(cols["cig_off"], np.uint32), cigar=_c(cols["cigar"], np.uint32))
    ncig = int(a["cig_off"][-1]) if n else 0
    cin = _CovIn(n, _p(a["tid"]), _p(a["pos"]), None, None, _p(a["cig_off"]), _p(a["cigar"]), 0, ncig)
    cap = cap_runs if cap_runs is not None else 2 * ncig + 16
    rt = np.zeros(cap, np.int32); rs = np.zeros(cap, np.int32); re = np.zeros(cap, np.int32); rv = np.zeros(cap, np.uint64)
    nout, bad = C.c_int64(0), C.c_int64(-1)
    f = lib().tbo_sample_heatmap
    f.restype = C.c_int
    vp = lambda x: C.c_void_p(_p(x))   # bare pointer arguments: ctypes would truncate plain ints to 32 bits
    rc = f(C.byref(cin), vp(a["yx"]), C.c_int64(cap), C.byref(nout), vp(rt), vp(rs), vp(re), vp(rv), C.byref(bad))
    if rc == 2:
        raise ValueError(f"unsupported CIGAR op in record {bad.value}")
    if rc != 0:
        raise RuntimeError(f"tbo_sample_heatmap rc={rc}")
    r = nout.value
    return rt[:r].copy(), rs[:r].copy(), re[:r].copy(), rv[:r].copy()


def heatmap_value(ival, n_samples):
    """normalize(bsam, 0.1, 1.5, n_samples) of src/tiecov.cpp:311-318 in float32, as printed with %f."""
    mult = np.float32(np.float32(1.5) - np.float32(0.1))
    return np.float32(np.float32(np.float32(ival) / np.float32(n_samples)) * mult + np.float32(0.1))
