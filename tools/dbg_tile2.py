"""Debug: generation-2 tile kernel vs generation 1 on the fixture / synthetic windows (YD on the sequential path so a bad
group table shows up as a diff, not as a YD failure)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["TB_YD_PATH"] = "seq"
import helpers as H
from tiebrush_b200 import api, sam, synth

def run(cols, run_off, gen, mode=0):
    os.environ["TB_TILE_GEN"] = str(gen)
    k = len(run_off) - 1
    with api.Context(device=0, n_samples=k, mode=mode) as ctx:
        try:
            r = ctx.collapse_window(cols, run_off)
        except Exception as ex:
            print("  gen", gen, "FAILED:", ex); return None
        print("  gen", gen, "tile_gen", ctx.last_tile_gen(), "heavy", ctx.last_heavy_slots(), ctx.last_tile_stats(), "G", r["n_groups"], "kept", r["n_kept"])
        return r

def cmp(a, b, cols):
    if a is None or b is None: return
    if a["n_groups"] != b["n_groups"]: print("  n_groups differ", a["n_groups"], b["n_groups"])
    g = min(a["n_groups"], b["n_groups"])
    for key in ("rep_index", "yc", "yx", "yd"):
        x, y = np.asarray(a[key])[:g], np.asarray(b[key])[:g]
        bad = np.nonzero(x != y)[0]
        if len(bad):
            i = int(bad[0])
            print(f"  {key}: {len(bad)} differ, first at group {i}: gen2 {x[i]} gen1 {y[i]}; pos gen2 {cols['pos'][int(np.asarray(a['rep_index'])[i])]} gen1 {cols['pos'][int(np.asarray(b['rep_index'])[i])]}")
            lo = max(0, i - 3)
            print("   gen2 reps", np.asarray(a["rep_index"])[lo:i + 4], "pos", cols["pos"][np.asarray(a["rep_index"])[lo:i + 4].astype(np.int64)])
            print("   gen1 reps", np.asarray(b["rep_index"])[lo:i + 4], "pos", cols["pos"][np.asarray(b["rep_index"])[lo:i + 4].astype(np.int64)])
        else:
            print(f"  {key}: equal")

for case in H.case_names("collapse_fixture.npz")[:2]:
    files, opts, fm, exp = H.load_collapse_case("collapse_fixture.npz", case)
    if opts.get("collapse_same"): continue
    for tid, cols, run_off, _src in sam.split_windows_by_tid(files):
        print("fixture", case, "tid", tid, "n", len(cols["pos"]), "k", len(run_off) - 1)
        cmp(run(cols, run_off, 2), run(cols, run_off, 1), cols)
for (n_tx, k, reads) in [(40, 12, 20000), (3, 5, 30000), (2000, 40, 5000)]:
    c, run_off, pr = synth.cohort_window(k, reads, seed=3, n_tx=n_tx, device="cpu")
    host = synth.to_host(c)
    print("synthetic", n_tx, k, reads)
    cmp(run(host, run_off, 2), run(host, run_off, 1), host)
