"""SAM-text record model and SoA window packer (host side, Python mirror).

Mirrors what the reference's record layer hands to the hot loops:
  * GSamRecord tag access      src/GSam.cpp:419-462  (tag_int / tag_float / tag_str)
  * GSamRecord::spliceStrand   src/GSam.cpp:464-475  (XS first, then minimap2 `ts` flipped on reverse)
  * passes_options' NH lookup  src/tiebrush.cpp:537   (NH absent => 0)
and lays the records out as the file-major struct-of-arrays window of include/tiebrush_b200.h.

This module is plumbing used by the tests, the golden generator and the Python CLI mirror; the
production packer that reads BAM directly is tiebrush_b200/csrc/host (C++).
"""
from __future__ import annotations

import hashlib
import numpy as np

CIGAR_OPS = "MIDNSHP=XB"
_OP_CODE = {c: i for i, c in enumerate(CIGAR_OPS)}


def parse_cigar(s: str) -> list[int]:
    """SAM CIGAR string -> packed BAM words (len<<4|op)."""
    if s == "*":
        return []
    out, num = [], 0
    for ch in s:
        if ch.isdigit():
            num = num * 10 + (ord(ch) - 48)
        else:
            out.append((num << 4) | _OP_CODE[ch])
            num = 0
    return out


def cigar_str(words) -> str:
    if len(words) == 0:
        return "*"
    return "".join(f"{int(w) >> 4}{CIGAR_OPS[int(w) & 0xF]}" for w in words)


def splice_strand(tags: dict, flag: int) -> str:
    """GSamRecord::spliceStrand (src/GSam.cpp:464-475) on SAM-text tags {name: (type, value)}."""
    c = ""
    xs = tags.get("XS")
    if xs is not None and xs[0] in ("A", "Z") and len(xs[1]) > 0:
        c = xs[1][0]
    if c == "":
        ts = tags.get("ts")
        m = ts[1][0] if (ts is not None and ts[0] in ("A", "Z") and len(ts[1]) > 0) else ""
        if m in "+-" and m != "":
            c = ({"+": "-", "-": "+"}[m]) if (flag & 16) else m
    return c if c in ("+", "-") else "."


def qname_hash(q: str) -> int:
    return int.from_bytes(hashlib.blake2b(q.encode(), digest_size=8).digest(), "little")


def line_hash(fields: list[str]) -> int:
    """64-bit content hash of a SAM line with the YC/YX/YD tags removed (identifies a record's bytes)."""
    keep = fields[:11] + [t for t in fields[11:] if t[:2] not in ("YC", "YX", "YD")]
    return int.from_bytes(hashlib.blake2b("\t".join(keep).encode(), digest_size=8).digest(), "little")


class SamRecords:
    """Column store of parsed SAM records (one file / one stream)."""

    def __init__(self):
        self.tid, self.pos, self.flag, self.mapq, self.strand, self.nh = [], [], [], [], [], []
        self.cigars, self.md, self.qhash, self.lhash = [], [], [], []
        self.yc, self.yx, self.yd, self.has_yc = [], [], [], []

    def __len__(self):
        return len(self.pos)


def parse_sam(text: str, ref_ids: dict | None = None):
    """Parse SAM text (header optional). Returns (SamRecords, ref_names)."""
    names = []
    ids = dict(ref_ids) if ref_ids else {}
    R = SamRecords()
    for line in text.split("\n"):
        if not line:
            continue
        if line[0] == "@":
            if line.startswith("@SQ"):
                for f in line.split("\t")[1:]:
                    if f.startswith("SN:"):
                        ids.setdefault(f[3:], len(ids))
                        names.append(f[3:])
            continue
        f = line.split("\t")
        flag = int(f[1])
        tags = {}
        for t in f[11:]:
            tags[t[:2]] = (t[3], t[5:])
        if f[2] not in ids:
            ids[f[2]] = len(ids)
        R.tid.append(ids[f[2]] if f[2] != "*" else -1)
        R.pos.append(int(f[3]) - 1)
        R.flag.append(flag)
        R.mapq.append(int(f[4]))
        R.strand.append(ord(splice_strand(tags, flag)))
        nh = tags.get("NH")
        nhv = int(nh[1]) if (nh is not None and nh[0] == "i") else 0
        R.nh.append(min(max(nhv, 0), 65535))
        R.cigars.append(parse_cigar(f[5]))
        md = tags.get("MD")
        R.md.append(md[1].encode() + b"\0" if (md is not None and md[0] in ("Z", "H")) else b"")
        R.qhash.append(qname_hash(f[0]))
        R.lhash.append(line_hash(f))
        yc = tags.get("YC")
        R.has_yc.append(yc is not None)
        R.yc.append(float(yc[1]) if (yc is not None and yc[0] in ("f", "i")) else 0.0)
        yx = tags.get("YX")
        R.yx.append(int(yx[1]) if (yx is not None and yx[0] == "i") else 1)
        yd = tags.get("YD")
        R.yd.append(int(yd[1]) if (yd is not None and yd[0] == "i") else 0)
    return R, ids


def to_columns(R: SamRecords) -> dict:
    """SamRecords -> dict of numpy columns (+ CSR cigar / md arenas)."""
    n = len(R)
    cig_off = np.zeros(n + 1, np.uint32)
    md_off = np.zeros(n + 1, np.uint32)
    if n:
        cig_off[1:] = np.cumsum([len(c) for c in R.cigars], dtype=np.int64).astype(np.uint32)
        md_off[1:] = np.cumsum([len(m) for m in R.md], dtype=np.int64).astype(np.uint32)
    cigar = np.fromiter((w for c in R.cigars for w in c), dtype=np.uint32, count=int(cig_off[-1]))
    md = np.frombuffer(b"".join(R.md), dtype=np.uint8).copy()
    return dict(
        tid=np.asarray(R.tid, np.int32), pos=np.asarray(R.pos, np.int32), flag=np.asarray(R.flag, np.uint16),
        mapq=np.asarray(R.mapq, np.uint8), strand=np.asarray(R.strand, np.uint8), nh=np.asarray(R.nh, np.uint16),
        cig_off=cig_off, cigar=cigar, md_off=md_off, md=md,
        qhash=np.asarray(R.qhash, np.uint64), lhash=np.asarray(R.lhash, np.uint64),
        yc_in=np.asarray(R.yc, np.float32), yx_in=np.asarray(R.yx, np.int32), yd_in=np.asarray(R.yd, np.int32),
        has_yc=np.asarray(R.has_yc, np.bool_),
    )


def _gather_csr(off, arena, idx):
    """Gather CSR rows `idx` -> (new_off, new_arena)."""
    lens = (off[1:] - off[:-1]).astype(np.int64)[idx]
    new_off = np.zeros(len(idx) + 1, np.uint32)
    new_off[1:] = np.cumsum(lens).astype(np.uint32)
    total = int(new_off[-1])
    if total == 0:
        return new_off, arena[:0].copy()
    starts = off[:-1].astype(np.int64)[idx]
    rep = np.repeat(starts - new_off[:-1].astype(np.int64), lens)
    src = rep + np.arange(total, dtype=np.int64)
    return new_off, arena[src]


PER_RECORD = ("tid", "pos", "flag", "mapq", "strand", "nh", "qhash", "lhash", "yc_in", "yx_in", "yd_in", "has_yc")


def take(cols: dict, idx) -> dict:
    """Row-subset of a column dict (CSR arenas re-packed)."""
    idx = np.asarray(idx, np.int64)
    out = {k: cols[k][idx] for k in PER_RECORD if k in cols}
    out["cig_off"], out["cigar"] = _gather_csr(cols["cig_off"], cols["cigar"], idx)
    if "md_off" in cols:
        out["md_off"], out["md"] = _gather_csr(cols["md_off"], cols["md"], idx)
    return out


def concat(parts: list[dict]) -> dict:
    """Concatenate column dicts (file-major window assembly)."""
    out = {}
    for k in PER_RECORD:
        if all(k in p for p in parts):
            out[k] = np.concatenate([p[k] for p in parts]) if parts else np.zeros(0)
    for offk, ak in (("cig_off", "cigar"), ("md_off", "md")):
        if not all(offk in p for p in parts):
            continue
        offs, base = [np.zeros(1, np.uint32)], 0
        for p in parts:
            offs.append((p[offk][1:].astype(np.int64) + base).astype(np.uint32))
            base += int(p[offk][-1])
        out[offk] = np.concatenate(offs)
        out[ak] = np.concatenate([p[ak] for p in parts]) if parts else np.zeros(0, np.uint32)
    return out


def split_windows_by_tid(files: list[dict]):
    """files: per-input-file column dicts (coordinate sorted). Yields (tid, window_cols, run_off, src_index)
    with one file-major window per reference id, in tid order. `src_index[j]` = (file, row) of window row j.
    Unmapped records (tid<0 / flag 4) are left out: the reference drops them before the collapse
    unless -M, and with -M it aborts."""
    tids = sorted({int(t) for f in files for t in np.unique(f["tid"]) if t >= 0})
    for tid in tids:
        parts, run_off, src = [], [0], []
        for fi, f in enumerate(files):
            idx = np.nonzero((f["tid"] == tid) & ((f["flag"] & 4) == 0))[0]
            parts.append(take(f, idx))
            run_off.append(run_off[-1] + len(idx))
            src.append(np.stack([np.full(len(idx), fi, np.int64), idx.astype(np.int64)], 1))
        yield tid, concat(parts), np.asarray(run_off, np.int64), np.concatenate(src) if src else np.zeros((0, 2), np.int64)
