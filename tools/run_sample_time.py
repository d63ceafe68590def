#!/usr/bin/env python
"""tiecov -s timing: device path (tc_sample_window, host columns) against the oracle (C port, 1 core) on one synthetic stream."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiebrush_b200 import api, synth
from oracle import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cols = synth.to_host(synth.coverage_stream(n, seed=5, chroms=1))
cols["yx"] = np.random.default_rng(5).integers(1, 61, size=n).astype(np.int32)
with api.Context(device=0, n_samples=1) as ctx:
    ctx.set_profiling(True)
    ctx.sample_window(cols)
    t0 = time.perf_counter(); got = ctx.sample_window(cols); dt = time.perf_counter() - t0
    kms = ctx.last_kernel_ms(1)
t0 = time.perf_counter(); exp = oracle.sample_heatmap(cols); dto = time.perf_counter() - t0
ok = all(np.array_equal(np.asarray(a), b) for a, b in zip(got, exp))
print(json.dumps({"records": n, "rows": int(len(exp[0])), "equal_to_oracle": bool(ok), "device_call_s": dt, "cell_kernel_ms": kms,
                  "oracle_s_1core": dto, "bases": int(synth.m_bases({"cigar": __import__("torch").as_tensor(cols["cigar"].view(np.int32))}))}))
