#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ by running the UNMODIFIED compiled
reference (oracle/_ref/{tiebrush,tiecov,htsfile}, built by oracle/build_ref.sh from /root/reference).

Run in the build container only (needs /root/reference for the fixture slices and oracle/_ref):
    python tests/golden/make_golden.py
Outputs (all small, committed):
    collapse_random.npz   randomized multi-file inputs x option sets  -> reference tiebrush output
    collapse_fixture.npz  coordinate slices of the reference's own test/t1 + test/t2 sample BAMs
    collapse_merged.npz   re-collapse of TieBrush-made inputs (tbMerged path)
    coverage.npz          tiecov -c/-j on collapsed and raw inputs     -> reference bedgraph / BED
Every case stores the parsed input columns (tiebrush_b200.sam layout) and the reference's answer.
"""
from __future__ import annotations

import os
import random
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tiebrush_b200 import sam  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
FIX = "/root/reference/test"
HEADER = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:chr1\tLN:1000000\n@SQ\tSN:chr2\tLN:500000\n"
REFIDS = {"chr1": 0, "chr2": 1}

# option sets: (name, argv, kwargs for the C-ABI)
KS, K2 = 1, 2
OPTSETS = [
    ("default", [], dict()),
    ("L", ["-L"], dict(mode=1)),
    ("P", ["-P"], dict(mode=2)),
    ("E", ["-E"], dict(mode=3)),
    ("F16", ["-F", "16"], dict(flag_mask=16)),
    ("F1040E", ["-F", "1040", "-E"], dict(flag_mask=1040, mode=3)),
    ("F83L", ["-F", "83", "-L"], dict(flag_mask=83, mode=1)),
    ("F3P", ["-F", "3", "-P"], dict(flag_mask=3, mode=2)),
    ("N1", ["-N", "1"], dict(max_nh=1)),
    ("Q30", ["-Q", "30"], dict(min_qual=30)),
    ("N2Q1S", ["-N", "2", "-Q", "1", "-S"], dict(max_nh=2, min_qual=1, keep_bits=KS)),
    ("K2SE", ["--keep-secondary", "-S", "-E"], dict(keep_bits=KS | K2, mode=3)),
    ("F2048SK2", ["-F", "2048", "-S", "--keep-secondary"], dict(flag_mask=2048, keep_bits=KS | K2)),
    ("A", ["-A"], dict(collapse_same=1)),
    ("SF", ["--store-frac", "--keep-secondary"], dict(keep_bits=K2 | 8)),
    ("SFF16", ["--store-frac", "--keep-secondary", "-F", "16"], dict(keep_bits=K2 | 8, flag_mask=16)),
]


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw)


def rand_cigar(rng: random.Random, allow_hx: bool):
    """A plausible CIGAR: [H][S] M (I|D|N M)* [S][H]; =/X occasionally when allowed."""
    ops = []
    if allow_hx and rng.random() < 0.05:
        ops.append((rng.randint(1, 5), "H"))
    if rng.random() < 0.15:
        ops.append((rng.randint(1, 6), "S"))
    nblocks = rng.choice([1, 1, 1, 2, 2, 3])
    for b in range(nblocks):
        m = "M" if not (allow_hx and rng.random() < 0.05) else rng.choice(["=", "X"])
        ops.append((rng.randint(3, 12), m))
        if b < nblocks - 1:
            r = rng.random()
            if r < 0.5:
                ops.append((rng.choice([20, 20, 35, 50]), "N"))
            elif r < 0.75:
                ops.append((rng.randint(1, 2), "D"))
            else:
                ops.append((rng.randint(1, 2), "I"))
    if rng.random() < 0.15:
        ops.append((rng.randint(1, 6), "S"))
    if allow_hx and rng.random() < 0.05:
        ops.append((rng.randint(1, 5), "H"))
    return "".join(f"{n}{o}" for n, o in ops)


def rand_files(seed: int, allow_hx=True, recycle_qnames=False):
    """2-6 coordinate-sorted SAM texts with heavy duplication across and inside files."""
    rng = random.Random(seed)
    nfiles = rng.randint(2, 6)
    # shared pools so duplicates are common
    positions = sorted(rng.sample(range(100, 400), 25)) + sorted(rng.sample(range(2000, 2200), 8))
    pool = {}
    for chrom in ("chr1", "chr2"):
        for p in positions:
            pool[(chrom, p)] = [rand_cigar(rng, allow_hx) for _ in range(rng.randint(1, 4))]
    mds = ["10", "5A4", "3^AC7", None]
    texts = []
    for fi in range(nfiles):
        recs = []
        n = rng.randint(150, 400)
        for i in range(n):
            chrom = "chr1" if rng.random() < 0.7 else "chr2"
            p = rng.choice(positions)
            cig = rng.choice(pool[(chrom, p)]) if rng.random() < 0.9 else rand_cigar(rng, allow_hx)
            flag = 0
            if rng.random() < 0.5:
                flag |= 16
            r = rng.random()
            if r < 0.3:
                flag |= 1 | (0x40 if rng.random() < 0.5 else 0x80) | (2 if rng.random() < 0.7 else 0) | (0x20 if rng.random() < 0.5 else 0)
            if rng.random() < 0.06:
                flag |= 0x100
            if rng.random() < 0.06:
                flag |= 0x800
            if rng.random() < 0.05:
                flag |= 0x400
            mapq = rng.choice([0, 1, 30, 60, 60, 60])
            tags = []
            nh = rng.choice([None, 1, 1, 1, 2, 5])
            if nh is not None:
                tags.append(f"NH:i:{nh}")
            s = rng.random()
            if s < 0.45:
                tags.append("XS:A:" + rng.choice("+-"))
            elif s < 0.6:
                tags.append("ts:A:" + rng.choice("+-"))
            elif s < 0.65:
                tags.append("XS:A:?")
            md = rng.choice(mds)
            if md is not None:
                tags.append("MD:Z:" + md)
            q = f"f{fi}.r{i}" if not recycle_qnames else f"q{rng.randint(0, 40)}"
            recs.append((REFIDS[chrom], p, i, "\t".join([q, str(flag), chrom, str(p), str(mapq), cig, "*", "0", "0", "*", "*"] + tags)))
        recs.sort(key=lambda t: (t[0], t[1], t[2]))
        body = "\n".join(r[3] for r in recs) + "\n"
        # a few unmapped reads at the tail (dropped by default; they must not disturb the merge)
        if rng.random() < 0.5:
            body += f"u{fi}\t4\t*\t0\t0\t*\t*\t0\t0\t*\t*\n"
        texts.append(HEADER + body)
    return texts


def ref_tiebrush(texts_or_paths, argv, tmp):
    paths = []
    for i, t in enumerate(texts_or_paths):
        if os.path.exists(t):
            paths.append(t)
        else:
            p = os.path.join(tmp, f"in{i}.sam")
            with open(p, "w") as f:
                f.write(t)
            paths.append(p)
    out = os.path.join(tmp, "out.bam")
    r = run([os.path.join(REF, "tiebrush")] + argv + ["-o", out] + paths)
    words = r.stderr.split()
    n_kept = int(words[0])
    n_out = int(words[words.index("as") + 1])
    txt = run([os.path.join(REF, "htsfile"), "-c", out]).stdout
    return txt, n_kept, n_out, out


def ref_tiecov(path, tmp):
    c, j = os.path.join(tmp, "k.cov"), os.path.join(tmp, "k.j")
    run([os.path.join(REF, "tiecov"), "-c", c, "-j", j, path])
    return open(c + ".bedgraph").read(), open(j + ".bed").read()


def pack_input(store: dict, inkey: str, file_cols: list[dict], with_tags=False):
    """Store the per-file input columns once; cases refer to them by key + file subset."""
    allc = sam.concat(file_cols)
    foff = np.cumsum([0] + [len(c["pos"]) for c in file_cols]).astype(np.int64)
    keys = ["tid", "pos", "flag", "mapq", "strand", "nh", "cig_off", "cigar", "md_off", "md", "lhash", "qhash"]
    if with_tags:
        keys += ["yc_in", "yx_in", "yd_in"]
    for k in keys:
        store[f"in/{inkey}/{k}"] = allc[k]
    store[f"in/{inkey}/file_off"] = foff


def pack_case(store: dict, key: str, inkey: str, files, out_cols: dict, n_kept: int, opts: dict, file_merged=None):
    store[f"case/{key}/inkey"] = np.asarray([inkey])
    store[f"case/{key}/files"] = np.asarray(list(files), np.int64)
    if file_merged is not None:
        store[f"case/{key}/file_merged"] = np.asarray(file_merged, np.uint8)
    store[f"case/{key}/opts"] = np.asarray([opts.get("mode", 0), opts.get("flag_mask", 0), opts.get("max_nh", 0x7FFFFFFF),
                                            opts.get("min_qual", -1), opts.get("keep_bits", 0), opts.get("collapse_same", 0)], np.int64)
    store[f"case/{key}/out/tid"] = out_cols["tid"]
    store[f"case/{key}/out/lhash"] = out_cols["lhash"]
    store[f"case/{key}/out/yc"] = out_cols["yc_in"]
    store[f"case/{key}/out/yx"] = out_cols["yx_in"]
    store[f"case/{key}/out/yd"] = out_cols["yd_in"]
    store[f"case/{key}/out/n_kept"] = np.asarray([n_kept], np.int64)


def parse_bedgraph(txt, ids):
    t, s, e, v = [], [], [], []
    for line in txt.split("\n"):
        if not line or line.startswith("track"):
            continue
        f = line.split("\t")
        t.append(ids[f[0]]); s.append(int(f[1])); e.append(int(f[2])); v.append(round(float(f[3]) * 1000))
    return np.asarray(t, np.int32), np.asarray(s, np.int32), np.asarray(e, np.int32), np.asarray(v, np.int64)


def parse_bed(txt, ids):
    t, s, e, n, v, c = [], [], [], [], [], []
    for line in txt.split("\n"):
        if not line or line.startswith("track"):
            continue
        f = line.split("\t")
        t.append(ids[f[0]]); s.append(int(f[1])); e.append(int(f[2])); n.append(int(f[3][4:]))
        v.append(round(float(f[4]) * 1000)); c.append(ord(f[5]))
    return (np.asarray(t, np.int32), np.asarray(s, np.int32), np.asarray(e, np.int32), np.asarray(n, np.int32),
            np.asarray(v, np.int64), np.asarray(c, np.uint8))


def pack_cov_case(store, key, in_cols, bg, bed):
    for k in ("tid", "pos", "flag", "strand", "cig_off", "cigar"):
        store[f"{key}/in/{k}"] = in_cols[k]
    store[f"{key}/in/yc"] = np.where(in_cols["has_yc"], in_cols["yc_in"], np.float32(1.0)).astype(np.float32)
    for nme, arr in zip(("tid", "start0", "end0", "milli"), bg):
        store[f"{key}/runs/{nme}"] = arr
    for nme, arr in zip(("tid", "start0", "end", "num", "milli", "strand"), bed):
        store[f"{key}/juncs/{nme}"] = arr


def fixture_slice_texts(lo_hi_by_chrom, tmp):
    """Coordinate slices of the reference's 20 sample BAMs, re-emitted as SAM text (same records)."""
    texts = []
    for d, pre in (("t1", "t1s"), ("t2", "t2s")):
        for i in range(10):
            txt = run([os.path.join(REF, "htsfile"), "-c", f"{FIX}/{d}/{pre}{i}.bam"]).stdout
            keep = []
            for line in txt.split("\n"):
                if not line:
                    continue
                if line[0] == "@":
                    keep.append(line)
                    continue
                f = line.split("\t", 5)
                rng = lo_hi_by_chrom.get(f[2])
                if rng and rng[0] <= int(f[3]) < rng[1]:
                    keep.append(line)
            texts.append("\n".join(keep) + "\n")
    return texts


def main():
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    cov_store = {}
    # ---- randomized collapse cases ---------------------------------------------------------
    store = {}
    ncase = 0
    for seed in range(8):
        for recycle in (False, True):
            texts = rand_files(1000 + seed, allow_hx=True, recycle_qnames=recycle)
            inkey = f"s{seed}{'q' if recycle else ''}"
            pack_input(store, inkey, [sam.to_columns(sam.parse_sam(t, REFIDS)[0]) for t in texts])
            for name, argv, kw in OPTSETS:
                if recycle != (name == "A"):
                    continue
                with tempfile.TemporaryDirectory() as tmp:
                    txt, n_kept, n_out, outp = ref_tiebrush(texts, argv, tmp)
                    outR, _ = sam.parse_sam(txt, REFIDS)
                    assert len(outR) == n_out
                    pack_case(store, f"s{seed}_{name}", inkey, range(len(texts)), sam.to_columns(outR), n_kept, kw)
                    ncase += 1
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collapse_random.npz"), **store)
    print("collapse_random:", ncase, "cases")

    # ---- coverage on raw randomized inputs (M/I/D/N/S only) --------------------------------
    for seed in range(10):
        texts = rand_files(2000 + seed, allow_hx=False)
        with tempfile.TemporaryDirectory() as tmp:
            # collapsed stream (has YC) and a raw single file (no YC)
            txt, _, _, outp = ref_tiebrush(texts, [], tmp)
            bg, bed = ref_tiecov(outp, tmp)
            R, ids = sam.parse_sam(txt, REFIDS)
            pack_cov_case(cov_store, f"col{seed}", sam.to_columns(R), parse_bedgraph(bg, ids), parse_bed(bed, ids))
            p = os.path.join(tmp, "raw.sam")
            open(p, "w").write(texts[0])
            bg, bed = ref_tiecov(p, tmp)
            R, ids = sam.parse_sam(texts[0], REFIDS)
            pack_cov_case(cov_store, f"raw{seed}", sam.to_columns(R), parse_bedgraph(bg, ids), parse_bed(bed, ids))

    # ---- fixture slices (reference's own test data) ----------------------------------------
    store = {}
    with tempfile.TemporaryDirectory() as tmp:
        texts = fixture_slice_texts({"chr12": (98_593_000, 98_595_500), "chr8": (24_950_000, 24_951_200)}, tmp)
        hdr_ids = sam.parse_sam(texts[0])[1]
        pack_input(store, "fix", [sam.to_columns(sam.parse_sam(t, hdr_ids)[0]) for t in texts])
        for name, argv, kw, sel in (("t1", [], {}, range(0, 10)), ("t2", [], {}, range(10, 20)), ("t12", [], {}, range(0, 20)),
                                    ("t12A", ["-A"], dict(collapse_same=1), range(0, 20)),
                                    ("t12E", ["-E"], dict(mode=3), range(0, 20))):
            sub = [texts[i] for i in sel]
            with tempfile.TemporaryDirectory() as tmp2:
                txt, n_kept, n_out, outp = ref_tiebrush(sub, argv, tmp2)
                outR, _ = sam.parse_sam(txt, hdr_ids)
                pack_case(store, name, "fix", sel, sam.to_columns(outR), n_kept, kw)
                print("fixture", name, "kept", n_kept, "out", n_out)
                if name == "t12":
                    bg, bed = ref_tiecov(outp, tmp2)
                    pack_cov_case(cov_store, "fix_t12", sam.to_columns(outR), parse_bedgraph(bg, hdr_ids), parse_bed(bed, hdr_ids))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collapse_fixture.npz"), **store)

    # ---- re-collapse of TieBrush-made inputs (tbMerged) -------------------------------------
    store = {}
    for seed in range(6):
        ta = rand_files(3000 + 2 * seed)
        tb = rand_files(3001 + 2 * seed)
        with tempfile.TemporaryDirectory() as tmp:
            da, db, dc = (os.path.join(tmp, x) for x in "abc")
            for d in (da, db, dc):
                os.makedirs(d)
            txa, _, _, pa = ref_tiebrush(ta, [], da)
            txb, _, _, pb = ref_tiebrush(tb, ["-E"] if seed % 2 else [], db)
            # merged + merged, and merged + raw file
            for name, inputs, merged in (("mm", [pa, pb], [1, 1]), ("mr", [pa, os.path.join(db, "in0.sam")], [1, 0])):
                txt, n_kept, n_out, _ = ref_tiebrush(inputs, [], dc)
                outR, _ = sam.parse_sam(txt, REFIDS)
                cols = []
                for pth in inputs:
                    t = run([os.path.join(REF, "htsfile"), "-c", pth]).stdout if pth.endswith(".bam") else open(pth).read()
                    cols.append(sam.to_columns(sam.parse_sam(t, REFIDS)[0]))
                pack_input(store, f"s{seed}_{name}", cols, with_tags=True)
                pack_case(store, f"s{seed}_{name}", f"s{seed}_{name}", range(len(cols)), sam.to_columns(outR), n_kept, {}, file_merged=merged)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "collapse_merged.npz"), **store)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "coverage.npz"), **cov_store)
    for f in ("collapse_random", "collapse_fixture", "collapse_merged", "coverage"):
        print(f, os.path.getsize(os.path.join(ROOT, "tests", "golden", f + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
