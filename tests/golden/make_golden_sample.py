#!/usr/bin/env python
"""Golden vectors for tiecov -s (groundwork): the reference's own fixtures test/t{1,2}/t{1,2}.bam (TieBrush-made)
as columns, and the rows (columns 1-4) of test/t{1,2}/t{1,2}.sample.bedgraph, which the reference's tests pin. Run where
/root/reference and oracle/_ref exist:  python tests/golden/make_golden_sample.py"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tiebrush_b200 import sam
out = {}
for t in ("t1", "t2"):
    text = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "htsfile"), "-c", f"/root/reference/test/{t}/{t}.bam"], capture_output=True, text=True, check=True).stdout
    R, ids = sam.parse_sam(text)
    cols = sam.to_columns(R)
    for k in ("tid", "pos", "yx_in", "cig_off", "cigar"):
        out[f"{t}/in/{k}"] = cols[k]
    out[f"{t}/n_samples"] = np.asarray([sum(1 for ln in text.split("\n") if ln.startswith("@CO\tSAMPLE:"))], np.int32)
    names = {k: v for k, v in ids.items()}
    rows = [ln.split("\t") for ln in open(f"/root/reference/test/{t}/{t}.sample.bedgraph").read().split("\n") if ln and not ln.startswith("track")]
    out[f"{t}/out/tid"] = np.asarray([names[r[0]] for r in rows], np.int32)
    out[f"{t}/out/start"] = np.asarray([int(r[1]) for r in rows], np.int32)
    out[f"{t}/out/end"] = np.asarray([int(r[2]) for r in rows], np.int32)
    out[f"{t}/out/ival"] = np.asarray([int(r[3]) for r in rows], np.uint64)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sample_heatmap.npz"), **out)
print({k: v.shape for k, v in out.items()})
