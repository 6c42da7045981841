"""One meme_matesw_batch on rescue-shaped pairs (626 x 150, bwa-meme_b200/e2e.py: make_rescue_pairs) for ncu / timing.
env: N pairs (default 131072 = one launch slice)."""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import synth, index_build, capi, e2e

n = int(os.environ.get("N", "131072"))
tmp = tempfile.mkdtemp(prefix="meme_msw_")
prefix = os.path.join(tmp, "g.fa")
index_build.build_index(prefix, synth.make_genome(30000, seed=2), bits=14)
ix = capi.MemeIndex(prefix, mode=3)
rec = e2e.bench_matesw(ix, n=n, threads=len(os.sched_getaffinity(0)), reps=int(os.environ.get("REPS", "3")), oracle=ge.load_oracle())
print(json.dumps(rec))
ix.close()
