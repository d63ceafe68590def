"""Deterministic synthetic RNA-seq cohorts, produced directly as SoA columns.

Vectorised torch code that runs unchanged on CPU (small parity cases) and on the GPU (BASELINE-size
cohorts are generated straight into HBM; there is no network and 10^9 records cannot be staged as BAM).

Transcript model (seed 42, shared by all samples): T transcripts per chromosome, 1-12 exons, exon length
~ round(lognormal(5.0,0.6)) in [30,3000], intron ~ round(lognormal(7.5,1.2)) in [70,200000], strand +/-
with p=.5, expression weight Zipf(s=1.1). Sample s (seed 1000+s): weight x lognormal(0,.5); reads drawn
proportional to weight x length, uniform offset, 150 bp projected to the genome => CIGAR of M/N; p=.04
one 1-3 bp I or D in the first block; p=.08 a 1-20 bp soft clip at one end; NH 1 (p=.9) else 2..10;
XS for spliced reads and 70 % of unspliced; MAPQ 60 if NH==1 else 0/1; flag 0x10 with p=.5
(`paired=True` adds 0x1,0x40/0x80,0x20 bits so -F has something to bite on); MD "150" or one mismatch.
"""
from __future__ import annotations

import numpy as np
import torch

READ_LEN = 150
MAXBLK = 6                      # a 150 bp read over exons >= 30 bp touches at most 6 exons
NCOL = 1 + 3 + 2 * (MAXBLK - 1) + 1
CHR1_LEN = 248_956_422
# GRCh38 primary chromosome lengths (chr1..22, X, Y), as in the fixtures' @SQ lines
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
          135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
          46709983, 50818468, 156040895, 57227415]
OP_M, OP_I, OP_D, OP_N, OP_S = 0, 1, 2, 3, 4


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


class TranscriptModel:
    """Exon chains on one chromosome (all tensors on `device`)."""

    def __init__(self, n_tx=20000, chrom_len=CHR1_LEN, seed=42, device="cpu"):
        g = _gen(seed, device)
        dev = device
        n_ex = torch.randint(1, 13, (n_tx,), generator=g, device=dev)
        exlen = torch.exp(5.0 + 0.6 * torch.randn((n_tx, 12), generator=g, device=dev)).round().clamp(30, 3000).long()
        inlen = torch.exp(7.5 + 1.2 * torch.randn((n_tx, 12), generator=g, device=dev)).round().clamp(70, 200000).long()
        col = torch.arange(12, device=dev)[None, :]
        live = col < n_ex[:, None]
        exlen = exlen * live
        # guarantee an mRNA of at least READ_LEN+50
        short = (exlen.sum(1) < READ_LEN + 50)
        exlen[:, 0] = torch.where(short, exlen[:, 0] + READ_LEN + 50, exlen[:, 0])
        inlen = inlen * (col < (n_ex[:, None] - 1))
        span = exlen.sum(1) + inlen.sum(1)
        start = (torch.rand((n_tx,), generator=g, device=dev, dtype=torch.float64) * (chrom_len - span - 2).clamp(min=1).double()).long() + 1
        # exon j starts at start + sum_{i<j}(exlen_i + inlen_i)
        step = exlen + inlen
        ex_start = start[:, None] + torch.cumsum(step, 1) - step
        cum = torch.zeros((n_tx, 13), dtype=torch.long, device=dev)
        cum[:, 1:] = torch.cumsum(exlen, 1)
        self.n_tx, self.n_ex, self.ex_start, self.ex_len, self.cum = n_tx, n_ex, ex_start, exlen, cum
        self.mrna_len = cum[:, 12]
        self.strand = torch.where(torch.rand((n_tx,), generator=g, device=dev) < 0.5, ord("+"), ord("-")).to(torch.uint8)
        rank = torch.randperm(n_tx, generator=g, device=dev).double() + 1.0
        self.weight = rank.pow(-1.1)
        self.device = dev


def sample_reads(tm: TranscriptModel, n_reads: int, seed: int, paired=False, with_md=False, yc_zipf=False, chunk=None):
    """One sample's reads, coordinate-sorted. Returns a dict of torch columns (+ cig_off / cigar arena)."""
    dev = tm.device
    g = _gen(seed, dev)
    w = tm.weight * torch.exp(0.5 * torch.randn((tm.n_tx,), generator=g, device=dev, dtype=torch.float64))
    p = w * (tm.mrna_len - READ_LEN + 1).clamp(min=1).double()
    tx = torch.multinomial(p / p.sum(), n_reads, replacement=True, generator=g)
    off = (torch.rand((n_reads,), generator=g, device=dev, dtype=torch.float64) * (tm.mrna_len[tx] - READ_LEN + 1).double()).long()
    cum = tm.cum[tx]                                            # [n,13]
    lo = torch.maximum(cum[:, :-1], off[:, None])
    hi = torch.minimum(cum[:, 1:], (off + READ_LEN)[:, None])
    m = (hi - lo).clamp(min=0)                                  # [n,12] bases of the read in each exon
    e0 = (cum[:, 1:] <= off[:, None]).sum(1)                    # first exon touched
    pos = tm.ex_start[tx, e0] + (off - tm.cum[tx, e0])          # 1-based start
    n = n_reads
    idx = (e0[:, None] + torch.arange(MAXBLK, device=dev)[None, :]).clamp(max=11)
    blk = torch.gather(m, 1, idx) * ((e0[:, None] + torch.arange(MAXBLK, device=dev)[None, :]) <= 11)   # [n,MAXBLK]
    exs = tm.ex_start[tx]; exl = tm.ex_len[tx]
    nxt = (idx + 1).clamp(max=11)
    gap = torch.gather(exs, 1, nxt) - (torch.gather(exs, 1, idx) + torch.gather(exl, 1, idx))          # intron after block j
    nblk = (blk > 0).sum(1)
    # ---- edits: clips first, then the indel inside what is left of the first block ----
    r = torch.rand((n, 6), generator=g, device=dev)
    g2 = torch.rand((n, 6), generator=g, device=dev)
    clip = (r[:, 4] < 0.08)
    clen = (r[:, 5] * 20).long() + 1
    clip_left = g2[:, 0] < 0.5
    last = (nblk - 1).clamp(min=0)
    blast = torch.gather(blk, 1, last[:, None])[:, 0]
    can_l = clip & clip_left & (blk[:, 0] > clen + 12)
    can_r = clip & ~clip_left & (blast > clen + 12)
    pos = torch.where(can_l, pos + clen, pos)
    blk = blk.clone()
    blk[:, 0] = torch.where(can_l, blk[:, 0] - clen, blk[:, 0])
    blk.scatter_(1, last[:, None], (torch.gather(blk, 1, last[:, None])[:, 0] - torch.where(can_r, clen, 0))[:, None])
    b0 = blk[:, 0]
    indel = (r[:, 0] < 0.04) & (b0 >= 12)
    is_ins = r[:, 1] < 0.5
    ilen = (r[:, 2] * 3).long() + 1
    ix = 2 + (r[:, 3] * (b0 - 8).clamp(min=1).float()).long()
    ix = torch.minimum(ix, (b0 - ilen - 2).clamp(min=2))
    # ---- op matrix ----
    ops = torch.zeros((n, NCOL), dtype=torch.long, device=dev)
    ops[:, 0] = torch.where(can_l, (clen << 4) | OP_S, 0)
    first_a = torch.where(indel, ix, b0)
    ops[:, 1] = (first_a << 4) | OP_M
    ops[:, 2] = torch.where(indel, (ilen << 4) | torch.where(is_ins, OP_I, OP_D), 0)
    rest = torch.where(is_ins, b0 - ix - ilen, b0 - ix)
    ops[:, 3] = torch.where(indel, (rest << 4) | OP_M, 0)
    for j in range(1, MAXBLK):
        live = blk[:, j] > 0
        ops[:, 4 + 2 * (j - 1)] = torch.where(live, (gap[:, j - 1] << 4) | OP_N, 0)
        ops[:, 5 + 2 * (j - 1)] = torch.where(live, (blk[:, j] << 4) | OP_M, 0)
    ops[:, NCOL - 1] = torch.where(can_r, (clen << 4) | OP_S, 0)
    order = torch.argsort(pos, stable=True)            # coordinate-sort the sample (after the clips moved starts)
    ops, pos, tx, nblk = ops[order], pos[order], tx[order], nblk[order]
    valid = ops != 0
    cnt = valid.sum(1)
    cig_off = torch.zeros(n + 1, dtype=torch.long, device=dev)
    cig_off[1:] = torch.cumsum(cnt, 0)
    cigar = ops[valid].to(torch.int32)  # values < 2^31 (lengths < 2^27); reinterpretable as uint32 words
    # ---- tags / flags ----
    nh = torch.where(g2[:, 1] < 0.9, 1, 2 + (torch.log(1 - g2[:, 2].clamp(max=0.999999)) / np.log(0.6)).long().clamp(max=8))
    spliced = nblk > 1
    has_xs = spliced | (g2[:, 3] < 0.7)
    strand = torch.where(has_xs, tm.strand[tx], torch.tensor(ord("."), dtype=torch.uint8, device=dev))
    mapq = torch.where(nh == 1, 60, (g2[:, 4] < 0.5).long())
    flag = torch.where(g2[:, 5] < 0.5, 16, 0)
    if paired:
        g3 = torch.rand((n, 3), generator=g, device=dev)
        flag = flag | 1 | torch.where(g3[:, 0] < 0.5, 0x40, 0x80) | torch.where(g3[:, 1] < 0.5, 0x20, 0) | torch.where(g3[:, 2] < 0.9, 2, 0)
    out = dict(pos=(pos - 1).to(torch.int32), flag=flag.to(torch.int16), mapq=mapq.to(torch.uint8), strand=strand,
               nh=nh.to(torch.int16), cig_off=cig_off.to(torch.int32), cigar=cigar, n_cig=int(cig_off[-1]))
    if with_md:
        # "150\0" or one mismatch "<a>X<b>\0" with p=.03; fixed 4..8 byte strings built on the host side of the tensor
        g4 = torch.rand((n, 2), generator=g, device=dev)
        mm = g4[:, 0] < 0.03
        a = (g4[:, 1] * 149).long()
        out["md_mm"], out["md_a"] = mm, a
    if yc_zipf:
        u = torch.rand((n,), generator=g, device=dev, dtype=torch.float64)
        out["yc"] = torch.floor(u.pow(-1.0 / 1.2)).clamp(max=1e6).to(torch.float32)
    return out


def md_columns(cols):
    """Materialise the MD arena ("150\\0" or "<a>T<149-a>\\0") for -L mode from the generator's draws."""
    mm = cols["md_mm"].cpu().numpy(); a = cols["md_a"].cpu().numpy()
    strs = [(f"{int(x)}T{149 - int(x)}" if f else "150").encode() + b"\0" for f, x in zip(mm, a)]
    off = np.zeros(len(strs) + 1, np.uint32)
    off[1:] = np.cumsum([len(s) for s in strs])
    return off, np.frombuffer(b"".join(strs), np.uint8).copy()


def md_columns_torch(cols, chunk=32_000_000):
    """Same arena as md_columns, vectorised on the columns' device (bench sizes): returns (md_off u32-as-int32 [n+1], md u8, n_md)."""
    mm_all, a_all = cols["md_mm"], cols["md_a"]
    dev = mm_all.device
    n = mm_all.numel()
    lens_parts, byte_parts = [], []

    def digits(x):   # left-aligned decimal digits of 0..199: ([m,3] uint8 chars, [m] lengths)
        d = torch.stack([x // 100, (x // 10) % 10, x % 10], 1)
        ln = 1 + (x >= 10).to(torch.int64) + (x >= 100).to(torch.int64)
        shift = 3 - ln
        idx = (torch.arange(3, device=dev)[None, :] + shift[:, None]).clamp_(max=2)
        return (torch.gather(d, 1, idx) + 48).to(torch.uint8), ln

    for c0 in range(0, n, chunk):
        mm = mm_all[c0:c0 + chunk].to(torch.bool); a = a_all[c0:c0 + chunk].to(torch.int64)
        m = a.numel()
        da, la = digits(a); db, lb = digits(149 - a)
        row = torch.zeros((m, 8), dtype=torch.uint8, device=dev)
        col = torch.arange(8, device=dev)[None, :]
        # mismatch rows: <a> T <149-a> NUL
        for j in range(3):
            row = torch.where((col == j) & (j < la)[:, None], da[:, j:j + 1], row)
            row = torch.where((col == (la + 1 + j)[:, None]) & (j < lb)[:, None], db[:, j:j + 1], row)
        row = torch.where(col == la[:, None], torch.full_like(row, 84), row)   # 'T'
        ln = la + 1 + lb + 1
        plain = torch.tensor([49, 53, 48, 0, 0, 0, 0, 0], dtype=torch.uint8, device=dev)   # "150\0"
        row = torch.where(mm[:, None], row, plain[None, :].expand(m, 8))
        ln = torch.where(mm, ln, torch.full_like(ln, 4))
        byte_parts.append(row[col < ln[:, None]])
        lens_parts.append(ln)
    lens = torch.cat(lens_parts)
    off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(lens, 0)
    md = torch.cat(byte_parts)
    return off.to(torch.int32), md, int(off[-1])


def prefix_slice(cols, run_off, target):
    """Coordinate prefix of a (device-resident) window holding about `target` records: every file's records with pos < hi,
    gathered on the columns' device and returned as HOST columns + the new run offsets (CPU baseline / oracle checks)."""
    dev = cols["pos"].device
    n = int(cols["pos"].shape[0])
    a0, b0 = int(run_off[0]), int(run_off[1])
    frac = min(1.0, target / max(n, 1))
    hi = int(cols["pos"][a0 + min(b0 - a0 - 1, int((b0 - a0) * frac))].item()) if frac < 1.0 else int(cols["pos"].max().item()) + 1
    parts, new_off = [], [0]
    for f in range(len(run_off) - 1):
        a, b = int(run_off[f]), int(run_off[f + 1])
        c = a + int(torch.searchsorted(cols["pos"][a:b], torch.tensor([hi], device=dev, dtype=cols["pos"].dtype))[0])
        parts.append(torch.arange(a, c, device=dev))
        new_off.append(new_off[-1] + (c - a))
    idx = torch.cat(parts)
    sub = {k: cols[k][idx].cpu().numpy() for k in ("pos", "flag", "mapq", "strand", "nh")}
    sub["flag"] = sub["flag"].view(np.uint16); sub["nh"] = sub["nh"].view(np.uint16)

    def gather_csr(off_key, arena_key):
        c0 = cols[off_key][idx].long() & 0xFFFFFFFF
        ln = (cols[off_key][idx + 1].long() & 0xFFFFFFFF) - c0
        off = torch.zeros(len(idx) + 1, dtype=torch.long, device=dev); off[1:] = torch.cumsum(ln, 0)
        rep = torch.repeat_interleave(torch.arange(len(idx), device=dev), ln)
        src = c0[rep] + (torch.arange(int(off[-1]), device=dev) - off[:-1][rep])
        return off.cpu().numpy().astype(np.uint32), cols[arena_key][src].cpu().numpy()

    sub["cig_off"], cig = gather_csr("cig_off", "cigar"); sub["cigar"] = cig.view(np.uint32)
    if "md_off" in cols:
        sub["md_off"], md = gather_csr("md_off", "md"); sub["md"] = md.view(np.uint8)
    return sub, np.asarray(new_off, np.int64)


def _cat_csr(parts, key_off="cig_off", key_arena="cigar"):
    offs, base = [torch.zeros(1, dtype=torch.int64, device=parts[0][key_off].device)], 0
    for p in parts:
        offs.append(p[key_off][1:].to(torch.int64) + base)
        base += int(p[key_off][-1])
    return torch.cat(offs).to(torch.int32), torch.cat([p[key_arena] for p in parts])


def cohort_window(n_samples, reads_per_sample, seed=0, n_tx=20000, chrom_len=CHR1_LEN, device="cpu", paired=False, with_md=False, tm=None):
    """A whole-chromosome, file-major collapse window: `n_samples` sorted runs on one tid.
    Returns (cols, run_off, pos_range)."""
    tm = tm or TranscriptModel(n_tx=n_tx, chrom_len=chrom_len, seed=42, device=device)
    parts = [sample_reads(tm, reads_per_sample, 1000 + seed * 100003 + s, paired=paired, with_md=with_md) for s in range(n_samples)]
    cols = {k: torch.cat([p[k] for p in parts]) for k in ("pos", "flag", "mapq", "strand", "nh")}
    cols["cig_off"], cols["cigar"] = _cat_csr(parts)
    cols["n_cig"] = int(sum(p["n_cig"] for p in parts))
    if with_md:
        cols["md_mm"] = torch.cat([p["md_mm"] for p in parts]); cols["md_a"] = torch.cat([p["md_a"] for p in parts])
    run_off = np.arange(n_samples + 1, dtype=np.int64) * reads_per_sample
    lo = int(cols["pos"].min()); hi = int(cols["pos"].max()) + 1
    return cols, run_off, (lo, hi)


def coverage_stream(n, seed=0, n_tx=20000, chroms=1, device="cpu", collapsed=True):
    """A coordinate-sorted (tid,pos) record stream for tiecov with integer YC weights (Zipf>=1 when `collapsed`)."""
    per = [n // chroms + (1 if c < n % chroms else 0) for c in range(chroms)]
    parts = []
    for c in range(chroms):
        clen = GRCH38[c % len(GRCH38)]
        tm = TranscriptModel(n_tx=max(64, n_tx // chroms), chrom_len=clen, seed=42 + c, device=device)
        p = sample_reads(tm, per[c], 5000 + seed * 7919 + c, yc_zipf=collapsed)
        p["tid"] = torch.full((per[c],), c, dtype=torch.int32, device=device)
        if not collapsed:
            p["yc"] = torch.ones(per[c], dtype=torch.float32, device=device)
        parts.append(p)
    cols = {k: torch.cat([p[k] for p in parts]) for k in ("tid", "pos", "yc", "strand")}
    cols["cig_off"], cols["cigar"] = _cat_csr(parts)
    cols["n_cig"] = int(sum(p["n_cig"] for p in parts))
    return cols


# ---- helpers used by tests / bench ------------------------------------------------------------------
def to_host(cols):
    out = {}
    for k, v in cols.items():
        if hasattr(v, "cpu"):
            a = v.cpu().numpy()
            if k in ("cig_off", "cigar"):
                a = a.view(np.uint32)
            elif k in ("flag", "nh"):
                a = a.view(np.uint16)
            out[k] = a
        else:
            out[k] = v
    return out


def ref_len_per_record(host):
    """Reference span and M-covered bases per record (host columns)."""
    cig = host["cigar"].astype(np.int64); off = host["cig_off"].astype(np.int64)
    op, ln = cig & 0xF, cig >> 4
    ref = np.where(np.isin(op, (0, 2, 3, 7, 8)), ln, 0)
    mcov = np.where(op == 0, ln, 0)
    cs_ref = np.concatenate([[0], np.cumsum(ref)]); cs_m = np.concatenate([[0], np.cumsum(mcov)])
    return cs_ref[off[1:]] - cs_ref[off[:-1]], cs_m[off[1:]] - cs_m[off[:-1]]


def covered_weight(cols):
    """sum over records of yc x (M bases) — equals sum over bedgraph runs of (end-start) x value."""
    host = to_host(cols)
    _, mcov = ref_len_per_record(host)
    return float((mcov.astype(np.float64) * host["yc"].astype(np.float64)).sum())


def bundle_cut(host, at_least):
    """Smallest record index >= at_least where a new tiecov bundle starts (tiecov.cpp:443), or n."""
    ref, _ = ref_len_per_record(host)
    end1 = host["pos"].astype(np.int64) + ref
    key = host["tid"].astype(np.int64) * (1 << 40) + end1
    pm = np.maximum.accumulate(key)
    head = np.ones(len(key), bool)
    head[1:] = (host["tid"][1:] != host["tid"][:-1]) | ((host["pos"][1:].astype(np.int64) + 1) > (pm[:-1] & ((1 << 40) - 1)))
    idx = np.nonzero(head[at_least:])[0]
    return int(at_least + idx[0]) if len(idx) else len(key)


def take_prefix(host, n):
    out = {k: host[k][:n] for k in ("tid", "pos", "yc", "strand") if k in host}
    out["cig_off"] = host["cig_off"][: n + 1].copy()
    out["cigar"] = host["cigar"][: int(out["cig_off"][-1])]
    return out


def m_bases(cols):
    """Sum of CIGAR M-op lengths ("coverage bases") of a torch column dict."""
    cig = cols["cigar"].to(torch.int64) & 0xFFFFFFFF
    return int(((cig >> 4) * ((cig & 0xF) == 0)).sum())


# ---- whole-genome collapsed stream for BASELINE config C4 (tiecov on 2e9 records), shardable by record index -------------
def end_column(cols, chunk=20_000_000):
    """0-based exclusive end (pos + reference length of the CIGAR: M D N = X) of every record of a torch column dict, int32,
    computed in chunks on the columns' device: GSamRecord::end as a host packer has it for free (GSam.cpp:351-417)."""
    n = int(cols["pos"].shape[0])
    dev = cols["pos"].device
    out = torch.empty(n, dtype=torch.int32, device=dev)
    off = cols["cig_off"]
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        o = off[a:b + 1].to(torch.int64) & 0xFFFFFFFF
        c0, c1 = int(o[0]), int(o[-1])
        cig = cols["cigar"][c0:c1].to(torch.int64) & 0xFFFFFFFF
        op = cig & 0xF
        ref = torch.where((op == 0) | (op == 2) | (op == 3) | (op == 7) | (op == 8), cig >> 4, torch.zeros_like(cig))
        cs = torch.zeros(c1 - c0 + 1, dtype=torch.int64, device=dev)
        cs[1:] = torch.cumsum(ref, 0)
        out[a:b] = (cols["pos"][a:b].to(torch.int64) + cs[o[1:] - c0] - cs[o[:-1] - c0]).to(torch.int32)
    return out


def genome_layout(n_total, split=4):
    """Contigs of the synthetic genome: every GRCh38 chromosome cut into `split` references (generation stays chunked:
    the largest holds 2 % of the records), records per contig proportional to its length. Returns (lengths, counts, starts)."""
    lens = np.repeat(np.asarray(GRCH38, np.int64) // split, split)
    w = lens / lens.sum()
    cnt = np.floor(w * n_total).astype(np.int64)
    cnt[: int(n_total - cnt.sum())] += 1
    start = np.concatenate([[0], np.cumsum(cnt)])
    return lens, cnt, start


def genome_slice(n_total, a, b, seed=0, n_tx=200000, device="cpu", seg_max=1_000_000_000, split=4, with_end=False):
    """Records [a, b) (global stream order) of the synthetic whole-genome collapsed stream of `n_total` records, as a list
    of device-resident segments (own CIGAR arena each, cut only at contig boundaries). The stream is the same whatever the
    slicing: contig c holds sample_reads(TranscriptModel(seed 42+c), counts[c], seed 5000+7919*seed+c) with Zipf YC."""
    lens, cnt, start = genome_layout(n_total, split)
    parts, segs, mb = [], [], 0

    def flush():
        if not parts:
            return
        cols = {k: torch.cat([p[k] for p in parts]) for k in ("tid", "pos", "yc", "strand")}
        cols["cig_off"], cols["cigar"] = _cat_csr(parts)
        cols["n_cig"] = int(sum(p["n_cig"] for p in parts))
        segs.append(cols)
        parts.clear()

    have = 0
    for c in range(len(cnt)):
        lo, hi = max(a, int(start[c])), min(b, int(start[c + 1]))
        if lo >= hi:
            continue
        tm = TranscriptModel(n_tx=max(64, int(n_tx * lens[c] / lens.sum())), chrom_len=int(lens[c]), seed=42 + c, device=device)
        p = sample_reads(tm, int(cnt[c]), 5000 + seed * 7919 + c, yc_zipf=True)
        i0, i1 = lo - int(start[c]), hi - int(start[c])
        off = p["cig_off"].to(torch.int64)
        c0, c1 = int(off[i0]), int(off[i1])
        q = dict(tid=torch.full((i1 - i0,), c, dtype=torch.int32, device=device), pos=p["pos"][i0:i1].clone(), yc=p["yc"][i0:i1].clone(),
                 strand=p["strand"][i0:i1].clone(), cig_off=(off[i0:i1 + 1] - c0).to(torch.int32), cigar=p["cigar"][c0:c1].clone(), n_cig=c1 - c0)
        del p, tm, off
        if have + (i1 - i0) > seg_max:
            flush(); have = 0
        parts.append(q); have += i1 - i0
    flush()
    for s in segs:
        mb += m_bases(s)
        if with_end:
            s["end"] = end_column(s)
    return segs, mb
