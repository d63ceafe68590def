"""Shared test plumbing: golden-case loading and window-by-window drivers for either implementation."""
from __future__ import annotations

import os
import numpy as np

from tiebrush_b200 import sam

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OPT_NAMES = ("mode", "flag_mask", "max_nh", "min_qual", "keep_bits", "collapse_same")
_cache = {}


def _npz(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return _cache[name]


def case_names(npz_name):
    z = _npz(npz_name)
    return sorted({k.split("/")[1] for k in z.files if k.startswith("case/")})


def _input_files(z, inkey, with_tags):
    pre = f"in/{inkey}/"
    cols = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    foff = cols.pop("file_off")
    n = len(cols["pos"])
    for k in ("yc_in", "yx_in", "yd_in"):
        if k not in cols:
            cols[k] = np.zeros(n, np.float32 if k == "yc_in" else np.int32)
    cols["has_yc"] = np.zeros(n, np.bool_)
    return [sam.take(cols, np.arange(foff[i], foff[i + 1])) for i in range(len(foff) - 1)]


def load_collapse_case(npz_name, case):
    z = _npz(npz_name)
    pre = f"case/{case}/"
    inkey = str(z[pre + "inkey"][0])
    files = _input_files(z, inkey, True)
    files = [files[i] for i in z[pre + "files"]]
    opts = dict(zip(OPT_NAMES, (int(v) for v in z[pre + "opts"])))
    fm = z[pre + "file_merged"] if (pre + "file_merged") in z.files else None
    exp = dict(tid=z[pre + "out/tid"], lhash=z[pre + "out/lhash"], yc=z[pre + "out/yc"], yx=z[pre + "out/yx"],
               yd=z[pre + "out/yd"], n_kept=int(z[pre + "out/n_kept"][0]))
    return files, opts, fm, exp


def run_collapse(collapse_fn, files, opts, file_merged=None):
    """Drive `collapse_fn(cols, run_off, tid=, file_merged=, **opts)` over one window per reference id."""
    o_tid, o_lh, o_yc, o_yx, o_yd, kept = [], [], [], [], [], 0
    for tid, cols, run_off, _src in sam.split_windows_by_tid(files):
        r = collapse_fn(cols, run_off, tid=tid, file_merged=file_merged, **opts)
        g = len(r["rep_index"])
        o_tid.append(np.full(g, tid, np.int32))
        o_lh.append(cols["lhash"][r["rep_index"].astype(np.int64)])
        o_yc.append(r["yc"]); o_yx.append(r["yx"]); o_yd.append(r["yd"])
        kept += r["n_kept"]
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return dict(tid=cat(o_tid, np.int32), lhash=cat(o_lh, np.uint64), yc=cat(o_yc, np.float32),
                yx=cat(o_yx, np.uint32), yd=cat(o_yd, np.int32), n_kept=kept)


def assert_collapse_equal(got, exp, what=""):
    assert got["n_kept"] == exp["n_kept"], f"{what}: n_kept {got['n_kept']} != {exp['n_kept']}"
    assert len(got["lhash"]) == len(exp["lhash"]), f"{what}: groups {len(got['lhash'])} != {len(exp['lhash'])}"
    for k in ("tid", "lhash", "yx", "yd"):
        bad = np.nonzero(got[k].astype(np.int64) != exp[k].astype(np.int64))[0] if k != "lhash" else np.nonzero(got[k] != exp[k])[0]
        assert len(bad) == 0, f"{what}: {k} differs at {bad[:5]} got {got[k][bad[:5]]} exp {exp[k][bad[:5]]}"
    # the golden YC went through htslib's SAM text ("%g", 6 significant digits): integers are exact, fractional values
    # (--store-frac) are compared after the same rendering
    gyc = np.asarray([np.float32(float("%g" % v)) for v in got["yc"]], np.float32) if len(got["yc"]) else np.zeros(0, np.float32)
    bad = np.nonzero(gyc.view(np.uint32) != exp["yc"].astype(np.float32).view(np.uint32))[0]
    assert len(bad) == 0, f"{what}: yc differs at {bad[:5]} got {got['yc'][bad[:5]]} exp {exp['yc'][bad[:5]]}"


def load_coverage_case(case):
    z = _npz("coverage.npz")
    pre = f"{case}/"
    cols = {k[len(pre) + 3:]: z[k] for k in z.files if k.startswith(pre + "in/")}
    keep = (cols["flag"] & 4) == 0  # tiecov.cpp:436-438
    n = len(cols["pos"])
    full = dict(cols)
    for k in sam.PER_RECORD:
        if k not in full:
            full[k] = np.zeros(n, np.int32)
    sub = sam.take({**full, "md_off": np.zeros(n + 1, np.uint32), "md": np.zeros(0, np.uint8)}, np.nonzero(keep)[0])
    sub["yc"] = cols["yc"][keep]
    runs = tuple(z[pre + "runs/" + k] for k in ("tid", "start0", "end0", "milli"))
    juncs = tuple(z[pre + "juncs/" + k] for k in ("tid", "start0", "end", "num", "milli", "strand"))
    return sub, runs, juncs


def coverage_case_names():
    z = _npz("coverage.npz")
    return sorted({k.split("/")[0] for k in z.files})


def milli(v):
    """'%.3f' rendering of a double as integer thousandths (exact for the integer-valued sums used here)."""
    return np.asarray([int(round(float(f"{x:.3f}") * 1000)) for x in v], np.int64)


def assert_coverage_equal(got, runs, juncs, what=""):
    gt, gs, ge, gv = got["runs"]
    assert len(gt) == len(runs[0]), f"{what}: runs {len(gt)} != {len(runs[0])}"
    assert (gt == runs[0]).all() and (gs == runs[1]).all() and (ge == runs[2]).all(), f"{what}: run coords differ"
    assert (milli(gv) == runs[3]).all(), f"{what}: run values differ"
    jt, js, je, jc, jv = got["juncs"]
    assert len(jt) == len(juncs[0]), f"{what}: juncs {len(jt)} != {len(juncs[0])}"
    assert (jt == juncs[0]).all() and (js - 1 == juncs[1]).all() and (je == juncs[2]).all(), f"{what}: junc coords differ"
    assert (np.arange(1, len(jt) + 1) == juncs[3]).all()
    assert (milli(jv) == juncs[4]).all() and (jc == juncs[5]).all(), f"{what}: junc values/strands differ"
