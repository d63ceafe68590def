"""End-to-end drop-in check of the host side (tiebrush_b200/host): the reference's own command lines with the hot loop
on the GPU (tiebrush_gpu / tiecov_gpu) against the UNMODIFIED reference binaries (oracle/_ref) on the same input files:
identical collapsed BAM records and tags, identical bedGraph and junction BED bytes. Inputs are synthetic SAM files of
the cohort model (the reference fixtures do not travel to the GPU box); all binaries are prebuilt by build()."""
import os
import subprocess

import numpy as np
import pytest

from tiebrush_b200 import sam, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
HOST = os.path.join(ROOT, "tiebrush_b200", "host", "_build")


def _need(*paths):
    missing = [p for p in paths if not os.path.exists(p)]
    if missing:
        pytest.skip("prebuilt binaries missing (built only where /root/reference exists): " + ", ".join(missing))


def _write_sams(tmp, k, reads, seed, n_tx, paired=False):
    cols, run_off, _ = synth.cohort_window(k, reads, seed=seed, n_tx=n_tx, device="cpu")
    host = synth.to_host(cols)
    rng = np.random.default_rng(seed)
    paths = []
    hdr = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:248956422\n"
    mds = ("150", "75A74", "10C139", None)
    for f in range(k):
        a, b = int(run_off[f]), int(run_off[f + 1])
        lines = [hdr]
        for i in range(a, b):
            cig = sam.cigar_str(host["cigar"][host["cig_off"][i]:host["cig_off"][i + 1]])
            flag = int(host["flag"][i])
            if paired:
                flag |= 0x1 | (0x40 if rng.random() < 0.5 else 0x80) | (0x20 if rng.random() < 0.5 else 0)
            if rng.random() < 0.03:
                flag |= 0x100
            tags = []
            nh = int(host["nh"][i])
            if nh:
                tags.append(f"NH:i:{nh}")
            s = chr(int(host["strand"][i]))
            if s != ".":
                tags.append(f"XS:A:{s}")
            md = mds[int(rng.integers(0, len(mds)))]
            if md is not None:
                tags.append(f"MD:Z:{md}")
            # recycled read names inside a file give -A something to look at
            name = f"s{f}.{int(rng.integers(0, max(2, (b - a) // 3)))}" if paired else f"s{f}.{i}"
            lines.append(f"{name}\t{flag}\tchr1\t{int(host['pos'][i]) + 1}\t{int(host['mapq'][i])}\t{cig}\t*\t0\t0\t*\t*\t" + "\t".join(tags) + "\n")
        p = os.path.join(tmp, f"s{f}.sam")
        with open(p, "w") as fh:
            fh.write("".join(lines))
        paths.append(p)
    return paths


def _records(bam):
    out = subprocess.run([os.path.join(REF, "htsfile"), "-c", bam], capture_output=True, text=True, check=True).stdout
    return [ln for ln in out.split("\n") if ln and not ln.startswith("@")]


def _run(cmd, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, f"{' '.join(cmd)} failed: {r.stderr[-500:]}"
    return r.stderr


@pytest.fixture(scope="module")
def sams(tmp_path_factory):
    _need(os.path.join(REF, "tiebrush"), os.path.join(REF, "htsfile"), os.path.join(HOST, "tiebrush_gpu"))
    tmp = str(tmp_path_factory.mktemp("cli"))
    pdir = os.path.join(tmp, "paired")
    os.makedirs(pdir, exist_ok=True)
    return tmp, _write_sams(tmp, k=6, reads=4000, seed=11, n_tx=25), _write_sams(pdir, k=5, reads=3000, seed=12, n_tx=12, paired=True)


@pytest.mark.parametrize("opts", [[], ["-E"], ["-P"], ["-L"], ["-N", "1", "-Q", "30"], ["-F", "16"], ["-S", "--keep-secondary", "-E"],
                                  ["--keep-secondary", "--store-frac"], ["-L", "-F", "272", "--keep-secondary"], ["-P", "-N", "2"],
                                  ["-E", "-Q", "1", "--keep-secondary"], ["-L", "-N", "5", "-Q", "1"], ["-P", "-F", "16", "-S"]])
def test_tiebrush_cli_matches_reference(sams, opts):
    tmp, paths, _ = sams
    tag = "_".join(o.strip("-") for o in opts) or "default"
    ref_out, our_out = os.path.join(tmp, f"ref_{tag}.bam"), os.path.join(tmp, f"gpu_{tag}.bam")
    ref_msg = _run([os.path.join(REF, "tiebrush")] + opts + ["-o", ref_out] + paths)
    our_msg = _run([os.path.join(HOST, "tiebrush_gpu")] + opts + ["-o", our_out] + paths)
    assert _records(our_out) == _records(ref_out)
    assert our_msg.strip().split("\n")[-1] == ref_msg.strip().split("\n")[-1]   # "N input records written as M (P% reduction)"


def test_tiebrush_cli_small_windows_and_collapse_same(sams):
    """Many small windows (cut at coverage gaps) must give the same bytes as one window; -A on recycled read names."""
    tmp, paths, paired = sams
    for name, files, opts, extra in (("w", paths, [], {}), ("a", paired, ["-A"], {}), ("af", paired, ["-A", "-F", "192"], {}),
                                     ("ww", paths, ["-E"], {"TB_WIRE_WIDE": "1"})):   # the last one: wide CIGAR columns instead of the compact wire format
        ref_out, our_out = os.path.join(tmp, f"ref_{name}.bam"), os.path.join(tmp, f"gpu_{name}.bam")
        _run([os.path.join(REF, "tiebrush")] + opts + ["-o", ref_out] + files)
        msg = _run([os.path.join(HOST, "tiebrush_gpu")] + opts + ["-o", our_out] + files,
                   env=dict({"TB_WINDOW_RECORDS": "300", "TB_WINDOW_SPAN": "2000", "TB_DECODE_THREADS": "3", "TB_TIMING": "1"}, **extra))
        assert _records(our_out) == _records(ref_out)
        assert "windows" in msg


def test_recollapse_and_tiecov_cli_match_reference(sams):
    """tiebrush -> tiebrush (TieBrush-made input) -> tiecov -c -j, every stage against the reference on the same file."""
    _need(os.path.join(REF, "tiecov"), os.path.join(HOST, "tiecov_gpu"))
    tmp, paths, _ = sams
    a, b = os.path.join(tmp, "half_a.bam"), os.path.join(tmp, "half_b.bam")
    _run([os.path.join(REF, "tiebrush"), "-o", a] + paths[:3])
    _run([os.path.join(REF, "tiebrush"), "-o", b] + paths[3:])
    ref_out, our_out = os.path.join(tmp, "ref_re.bam"), os.path.join(tmp, "gpu_re.bam")
    _run([os.path.join(REF, "tiebrush"), "-o", ref_out, a, b])
    _run([os.path.join(HOST, "tiebrush_gpu"), "-o", our_out, a, b])
    assert _records(our_out) == _records(ref_out)
    for src, tag in ((ref_out, "collapsed"), (paths[0], "raw")):
        rc, rj = os.path.join(tmp, f"ref_{tag}_cov"), os.path.join(tmp, f"ref_{tag}_junc")
        gc, gj = os.path.join(tmp, f"gpu_{tag}_cov"), os.path.join(tmp, f"gpu_{tag}_junc")
        _run([os.path.join(REF, "tiecov"), "-c", rc, "-j", rj, src])
        _run([os.path.join(HOST, "tiecov_gpu"), "-c", gc, "-j", gj, src], env={"TB_WINDOW_RECORDS": "500"})
        assert open(gc + ".bedgraph").read() == open(rc + ".bedgraph").read()
        assert open(gj + ".bed").read() == open(rj + ".bed").read()
        assert len(open(rc + ".bedgraph").read().split("\n")) > 10
    # -s (sample heat-map) needs the @CO SAMPLE lines of a TieBrush-made header: the collapsed files only
    for src, tag in ((ref_out, "re"), (a, "half")):
        rs, gs = os.path.join(tmp, f"ref_{tag}_heat"), os.path.join(tmp, f"gpu_{tag}_heat")
        _run([os.path.join(REF, "tiecov"), "-s", rs, src])
        _run([os.path.join(HOST, "tiecov_gpu"), "-s", gs, "-c", gs + "_c", src], env={"TB_WINDOW_RECORDS": "700"})
        assert open(gs + ".bedgraph").read() == open(rs + ".bedgraph").read()
        assert len(open(rs + ".bedgraph").read().split("\n")) > 10
