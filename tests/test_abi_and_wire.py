"""CPU-side checks (no GPU): the C-ABI library loads and exports every function include/tiebrush_b200.h declares; the host
packers of the wire formats decode back to the wide columns by the rules the device applies; the synthetic C4 stream is the
same whatever slice of it a rank draws."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """No compute call (no GPU here): dlopen + symbol lookup of everything the header declares."""
    from tiebrush_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "tiebrush_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(t[bc]_[a-z0-9_]+)\s*\(", hdr))
    assert {"tb_create", "tb_collapse_window", "tc_coverage_window", "tc_coverage_stream", "tc_shard_coverage_gather", "tb_comm_init"} <= declared
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.SYMBOLS) <= declared


def test_ctypes_structs_match_the_header_field_order():
    from tiebrush_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "tiebrush_b200.h")).read()
    for struct, cls in (("tb_soa_in", _lib.SoaIn), ("tc_soa_in", _lib.CovIn), ("tb_groups_out", _lib.GroupsOut), ("tc_runs_out", _lib.RunsOut), ("tc_juncs_out", _lib.JuncsOut)):
        body = hdr[:hdr.index("} " + struct + ";")]
        body = body[body.rindex("typedef struct {"):]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:,|;)", body)
        names = [n for n in names if n not in ("typedef", "struct")]
        assert names == [f[0] for f in cls._fields_], (struct, names)


def _decode_fixed(pk):
    d8, ext = pk["pos_d8"].astype(np.int64), pk["pos_ext"].astype(np.int64)
    esc = d8 >= 254
    val = d8.copy(); val[esc] = ext
    pos = np.zeros(len(d8), np.int64); cur = 0
    for i in range(len(d8)):
        cur = val[i] if d8[i] == 255 else cur + val[i]
        pos[i] = cur
    m8 = pk["meta8"]
    tup = np.zeros(len(m8), np.uint64)
    hit = m8 != 255
    tup[hit] = pk["meta_dict"][m8[hit]]
    tup[~hit] = pk["meta_ext"]
    return pos, (tup & np.uint64(0xffff)).astype(np.uint16), ((tup >> np.uint64(16)) & np.uint64(0xff)).astype(np.uint8), \
        ((tup >> np.uint64(24)) & np.uint64(0xff)).astype(np.uint8), ((tup >> np.uint64(32)) & np.uint64(0xffff)).astype(np.uint16)


@pytest.mark.parametrize("max_dict", [255, 2, 0])
def test_packed_fixed_columns_round_trip(max_dict):
    """api.pack_fixed_columns: pos_d8 / pos_ext / meta8 / meta_dict / meta_ext decode back to pos, flag, mapq, strand, nh by
    the rule of include/tiebrush_b200.h (the one the device applies), numpy and torch packers alike."""
    from tiebrush_b200 import api, synth
    cols, run_off, _ = synth.cohort_window(7, 2500, seed=4, n_tx=15, device="cpu", paired=True)
    host = synth.to_host(cols)
    host["pos"] = host["pos"].copy()
    host["pos"][100] += 5_000_000          # a gap far beyond 253 -> escape 254 ... and a record out of order behind it -> absolute
    pk = api.pack_fixed_columns(host, run_off, max_dict=max_dict)
    assert len(pk["meta_dict"]) <= max_dict and int((pk["pos_d8"] == 255).sum()) >= 7 and int((pk["pos_d8"] == 254).sum()) >= 1
    assert int((pk["pos_d8"] >= 254).sum()) == len(pk["pos_ext"]) and int((pk["meta8"] == 255).sum()) == len(pk["meta_ext"])
    pos, flag, mapq, strand, nh = _decode_fixed(pk)
    assert np.array_equal(pos, host["pos"]) and np.array_equal(flag, host["flag"]) and np.array_equal(mapq, host["mapq"])
    assert np.array_equal(strand, host["strand"]) and np.array_equal(nh, host["nh"])
    tcols = dict(cols); tcols["pos"] = cols["pos"].clone(); tcols["pos"][100] += 5_000_000
    tk = api.pack_fixed_columns(tcols, run_off, max_dict=max_dict)     # the torch packer may pick other dictionary entries among ties
    tk = {kk: (v.numpy() if hasattr(v, "numpy") else v) for kk, v in tk.items()}
    tk["meta_ext"] = tk["meta_ext"].view(np.uint64)
    assert np.array_equal(tk["pos_d8"], pk["pos_d8"]) and np.array_equal(tk["pos_ext"], pk["pos_ext"])
    pos, flag, mapq, strand, nh = _decode_fixed(tk)
    assert np.array_equal(pos, host["pos"]) and np.array_equal(flag, host["flag"]) and np.array_equal(mapq, host["mapq"])
    assert np.array_equal(strand, host["strand"]) and np.array_equal(nh, host["nh"])


def test_genome_slices_are_pieces_of_one_stream():
    """synth.genome_slice (BASELINE C4 stream): any [a, b) is the same records whatever else is drawn; end = pos + reference length."""
    import torch
    from tiebrush_b200 import synth
    R = 120000
    whole, mb = synth.genome_slice(R, 0, R, n_tx=3000, with_end=True)
    cat = {k: torch.cat([s[k] for s in whole]) for k in ("tid", "pos", "yc", "strand", "end")}
    assert int(cat["pos"].shape[0]) == R and bool((cat["tid"][1:] >= cat["tid"][:-1]).all())
    same_tid = cat["tid"][1:] == cat["tid"][:-1]
    assert bool((cat["pos"][1:][same_tid] >= cat["pos"][:-1][same_tid]).all())
    host = synth.to_host(whole[0])
    ref, _ = synth.ref_len_per_record(host)
    assert np.array_equal(host["end"], host["pos"] + ref)
    for a, b in ((0, 17), (41000, 97531), (R - 5, R)):
        part, _ = synth.genome_slice(R, a, b, n_tx=3000, with_end=True)
        for k in ("tid", "pos", "yc", "strand", "end"):
            assert torch.equal(torch.cat([s[k] for s in part]), cat[k][a:b]), (k, a, b)
    lens, cnt, start = synth.genome_layout(2_000_000_000)
    assert len(cnt) == 96 and int(cnt.sum()) == 2_000_000_000 and int(cnt.max()) < 45_000_000
