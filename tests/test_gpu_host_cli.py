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


# ---- BASELINE config C1: the reference's own 20 fixture BAMs through the GPU command lines --------------------------------
FIX = os.path.join(ROOT, "tests", "golden", "fixtures")     # test/t1/t1s*.bam and test/t2/t2s*.bam of the reference (data, 9.3 MB)
# record-text md5 (htsfile -c out.bam | grep -v '^@' | md5sum) and output-file md5s of the compiled reference, SURVEY §8c / BASELINE.md §3
C1_MD5 = {"t1": "db674977026be7f2832fe9869ed40376", "t2": "27f7b93a789a9cddb06dca63449d32af", "t12": "50320443e9d380da6bc6c46b1fa90d85",
          "cov": "8d1b0c70113a233095742cec61bb8129", "junc": "61427ff301f9ea70d48d5fbff20db7e2"}


def _fixture_bams(which):
    import glob
    fs = []
    for t in which:
        fs += sorted(glob.glob(os.path.join(FIX, t, f"{t}s[0-9].bam")))
    return fs


def _md5_records(bam):
    import hashlib
    return hashlib.md5(("\n".join(_records(bam)) + "\n").encode()).hexdigest()


@pytest.mark.parametrize("name,which", [("t1", ["t1"]), ("t2", ["t2"]), ("t12", ["t1", "t2"])])
def test_c1_fixture_bams_through_tiebrush_gpu_md5(tmp_path, name, which):
    """C1: `tiebrush -o o.bam test/t1/t1s[0-9].bam ...` on the GPU command line reproduces the record md5 of the compiled reference."""
    _need(os.path.join(REF, "htsfile"), os.path.join(HOST, "tiebrush_gpu"))
    files = _fixture_bams(which)
    assert len(files) == 10 * len(which)
    out = str(tmp_path / f"{name}.bam")
    msg = _run([os.path.join(HOST, "tiebrush_gpu"), "-o", out] + files)
    assert _md5_records(out) == C1_MD5[name]
    assert {"t1": "416922 input records written as 3479", "t2": "242910 input records written as 8179", "t12": "659832 input records written as 9491"}[name] in msg


def test_c1_tiecov_gpu_on_the_collapsed_fixtures_md5(tmp_path):
    """C1: tiecov -c -j on the 20-file collapsed BAM: bedGraph and junction BED bytes of the compiled reference (md5)."""
    import hashlib
    _need(os.path.join(HOST, "tiebrush_gpu"), os.path.join(HOST, "tiecov_gpu"))
    out = str(tmp_path / "t12.bam")
    _run([os.path.join(HOST, "tiebrush_gpu"), "-o", out] + _fixture_bams(["t1", "t2"]))
    c, j = str(tmp_path / "k.cov"), str(tmp_path / "k.j")
    _run([os.path.join(HOST, "tiecov_gpu"), "-c", c, "-j", j, out], env={"TB_WINDOW_RECORDS": "2000"})
    assert hashlib.md5(open(c + ".bedgraph", "rb").read()).hexdigest() == C1_MD5["cov"]
    assert hashlib.md5(open(j + ".bed", "rb").read()).hexdigest() == C1_MD5["junc"]


def test_store_frac_then_tiecov_cli_is_exact(sams):
    """tiebrush --store-frac writes YC like 0.333333: tiecov_gpu must print the reference's bytes (ordered double sums), not
    2^-20 fixed-point sums. VERDICT r1 item 5a."""
    _need(os.path.join(REF, "tiecov"), os.path.join(HOST, "tiecov_gpu"))
    tmp, paths, _ = sams
    frac = os.path.join(tmp, "frac.bam")
    _run([os.path.join(REF, "tiebrush"), "--store-frac", "--keep-secondary", "-o", frac] + paths)
    assert any("YC:f:0.3" in r or "YC:f:0.1" in r or "YC:f:0.2" in r for r in _records(frac)[:5000])
    rc, rj, gc, gj = (os.path.join(tmp, x) for x in ("ref_frac_cov", "ref_frac_j", "gpu_frac_cov", "gpu_frac_j"))
    _run([os.path.join(REF, "tiecov"), "-c", rc, "-j", rj, frac])
    _run([os.path.join(HOST, "tiecov_gpu"), "-c", gc, "-j", gj, frac], env={"TB_WINDOW_RECORDS": "800"})
    assert open(gc + ".bedgraph").read() == open(rc + ".bedgraph").read()
    assert open(gj + ".bed").read() == open(rj + ".bed").read()


def test_tiecov_cli_on_several_gpus_matches_reference(sams):
    """TB_DEVICES: tiecov_gpu shards every round of the stream over the GPUs of the box (tc_shard_coverage: NCCL halo exchange)
    and prints the reference's bytes, with rounds small enough that cuts fall inside bundles. Needs >= 2 GPUs."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    _need(os.path.join(REF, "tiecov"), os.path.join(REF, "tiebrush"), os.path.join(HOST, "tiecov_gpu"))
    tmp, paths, _ = sams
    merged = os.path.join(tmp, "multi_in.bam")
    _run([os.path.join(REF, "tiebrush"), "-o", merged] + paths)
    rc, rj = os.path.join(tmp, "ref_multi_cov"), os.path.join(tmp, "ref_multi_j")
    _run([os.path.join(REF, "tiecov"), "-c", rc, "-j", rj, merged])
    for devs, wr in ((f"0-{min(ngpu, 8) - 1}", "150"), ("0,1", "40")):
        gc, gj = os.path.join(tmp, "gpu_multi_cov" + wr), os.path.join(tmp, "gpu_multi_j" + wr)
        msg = _run([os.path.join(HOST, "tiecov_gpu"), "-c", gc, "-j", gj, merged], env={"TB_DEVICES": devs, "TB_WINDOW_RECORDS": wr, "TB_TIMING": "1"})
        assert "GPUs" in msg and "rounds" in msg
        assert open(gc + ".bedgraph").read() == open(rc + ".bedgraph").read()
        assert open(gj + ".bed").read() == open(rj + ".bed").read()


def test_tiebrush_cli_on_several_gpus_matches_reference(sams):
    """TB_DEVICES: tiebrush_gpu computes its windows on several GPUs at once and writes them in hand-over order. Needs >= 2 GPUs."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    _need(os.path.join(REF, "tiebrush"), os.path.join(REF, "htsfile"), os.path.join(HOST, "tiebrush_gpu"))
    tmp, paths, _ = sams
    for opts in ([], ["-E", "-N", "2"]):
        tag = "multi" + "_".join(o.strip("-") for o in opts)
        ref_out, our_out = os.path.join(tmp, f"ref_{tag}.bam"), os.path.join(tmp, f"gpu_{tag}.bam")
        ref_msg = _run([os.path.join(REF, "tiebrush")] + opts + ["-o", ref_out] + paths)
        our_msg = _run([os.path.join(HOST, "tiebrush_gpu")] + opts + ["-o", our_out] + paths,
                       env={"TB_DEVICES": f"0-{min(ngpu, 8) - 1}", "TB_WINDOW_RECORDS": "250", "TB_WINDOW_SPAN": "2000", "TB_TIMING": "1"})
        assert _records(our_out) == _records(ref_out)
        assert "GPU(s)" in our_msg and ref_msg.strip().split("\n")[-1] in our_msg
