import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.load_pkg(); O = ge.load_oracle()
from bwa_meme_b200 import capi, synth, index_build
d = "/tmp/meme_dbg"; os.makedirs(d, exist_ok=True)
prefix = d + "/g.fa"
g = synth.make_genome(60000, seed=5)
if not os.path.exists(prefix + ".pos_packed"):
    index_build.build_index(prefix, g, bits=12)
n = int(os.environ.get("N", "64")); rounds = int(os.environ.get("ROUNDS", "3"))
codes, offs = synth.make_reads(g, n, 150, seed=11, n_frac=0.05)
ix = capi.MemeIndex(prefix, mode=int(os.environ.get("MODE", "3")))
got = ix.seed(codes, offs, capi.default_seed_opt(rounds=rounds))
oix = O.OracleIndex(prefix, 3)
want = oix.seed(codes, offs, O.default_opt(rounds=rounds))
a = O.canon_lines(got["smems"], got["smem_off"], got["hit_off"], got["hits"], n)
b = O.canon_lines(want["smems"], want["smem_off"], want["hit_off"], want["hits"], n)
bad = [i for i in range(n) if a[i] != b[i]]
print("reads", n, "rounds", rounds, "mismatching", len(bad), ix.seed_counters())
for i in bad[:3]:
    print(i, "GPU", a[i][:8]); print(i, "ORA", b[i][:8])
