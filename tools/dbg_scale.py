import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tiebrush_b200 import api, synth
k = 100
reads = int(sys.argv[1])
dev = torch.device("cuda", 0)
cols, run_off, pr = synth.cohort_window(k, reads, seed=0, device=dev)
torch.cuda.synchronize()
print("generated", reads, "n_cig", cols["n_cig"], "mem", torch.cuda.memory_allocated() / 1e9, flush=True)
n = k * reads
ctx = api.Context(device=0, n_samples=k)
out = dict(rep_index=torch.empty(n, dtype=torch.int32, device=dev), yc=torch.empty(n, dtype=torch.float32, device=dev),
           yx=torch.empty(n, dtype=torch.int32, device=dev), yd=torch.empty(n, dtype=torch.int32, device=dev))
for it in range(2):
    t = time.time()
    try:
        r = ctx.collapse_window(cols, run_off, pos_range=pr, out=out)
        print("ok", r["n_groups"], r["n_kept"], "path", ctx.last_path(), ctx.last_yd_path(), time.time() - t, flush=True)
    except Exception as e:
        print("ERR", e, flush=True)
