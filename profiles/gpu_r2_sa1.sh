#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_index_gpu.py -x -q 2>&1 | tail -15
timeout 900 python - <<'P' 2>&1 | tail -12
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import synth, index_build, capi
for G in (20_000_000, 248_956_422):
    g = synth.make_genome_large(G, n_families=30, copies=100)
    text = index_build.fwd_rc_text(g)
    t0 = time.time(); r = capi.sa_build(text); t1 = time.time()
    print(f"native SA + keys, {len(text)} suffixes: {t1 - t0:.2f} s, {r['rounds']} doubling rounds, pad {r['pad']}", flush=True)
    t0 = time.time(); sa = index_build.suffix_array(text, device="cuda"); torch.cuda.synchronize(); t1 = time.time()
    same = bool((sa.cpu().numpy() == r["sa"]).all())
    print(f"torch prefix doubling: {t1 - t0:.2f} s; identical: {same}", flush=True)
    del sa, r; torch.cuda.empty_cache()
P
