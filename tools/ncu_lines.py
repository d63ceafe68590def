#!/usr/bin/env python
"""Per-source-line instruction counts from `ncu --page source --print-source cuda,sass --csv` output.
usage: ncu_lines.py file.csv [top_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # inst, samples, stall_barrier
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr, r))
    # columns 'Source' appear twice (cuda, sass): csv.reader keeps order, dict keeps the last (sass)
    try:
        inst = int(d.get("Instructions Executed", "0") or 0); samp = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    a = agg[(cur_file, line)]; a[0] += inst; a[1] += samp; a[3] = r[1]
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print(f"total warp-instructions {tot}  samples {tots}")
for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*a[0]/max(tot,1):5.1f}% inst {100*a[1]/max(tots,1):5.1f}% smp  {f}:{l}  {a[3].strip()[:110]}")
