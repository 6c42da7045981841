"""Banded-SW kernel micro-benchmark: extension-like pairs, HBM-resident, GCUPS + pairs/s."""
import sys, os, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg(); O = ge.load_oracle()
from bwa_meme_b200 import capi
import torch
n = int(os.environ.get("NP", "200000")); reps = int(os.environ.get("REPS", "8"))
pairs, ref, qer = O.make_pairs(n, seed=5, max_q=150, max_extra_ref=110, h0_max=120)
big = np.tile(pairs, reps)
b = capi.MemeBsw()
b.upload(big, ref, qer)
for _ in range(2): b.run(100)
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 5
for _ in range(K): b.run(100)
out = big.copy(); b.download(out)
dt = (time.perf_counter() - t0) / K
cells = b.cells()
want = O.bsw_batch(pairs[:5000], ref, qer, 100)
ok = all((out[f][:5000] == want[f]).all() for f in O.RESULT_FIELDS)
print(json.dumps({"pairs": len(big), "ms": dt * 1e3, "pairs_per_s": len(big) / dt, "gcups": cells / dt / 1e9, "parity_5000": bool(ok),
                  "mean_qlen": float(pairs["len2"].mean()), "mean_tlen": float(pairs["len1"].mean()), "cells_per_pair": cells / len(big)}))
