import sys, os, numpy as np, time
sys.path.insert(0,'/root/repo'); import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import capi, synth, index_build
import torch
d="/tmp/meme_prof"; os.makedirs(d,exist_ok=True)
prefix=d+"/ref.fa"
G=int(os.environ.get("G","20000000")); N=int(os.environ.get("N","400000"))
if not os.path.exists(prefix+".pos_packed"):
    g=synth.make_genome_large(G, n_families=30, copies=100)
    g.tofile(d+"/genome.u8")
    index_build.build_index(prefix,g,device="cuda",write_keys_file=False)
g=np.fromfile(d+"/genome.u8",dtype=np.uint8)
reads=synth.make_reads_torch(torch.from_numpy(g).cuda(), N, 150, seed=3).cpu().numpy()
offs=np.arange(N+1,dtype=np.int64)*150
ix=capi.MemeIndex(prefix,mode=3)
opt=capi.default_seed_opt()
ix.seed_upload(reads.reshape(-1),offs)
for i in range(3):
    ix.seed_run(opt); print(ix.seed_last_ms(), ix.seed_totals())
print(ix.seed_counters())
