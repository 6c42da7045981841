#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_index_gpu.py -x -q 2>&1 | tail -15
timeout 900 python - <<'P' 2>&1 | tail -12
import sys, time, json, numpy as np, torch
sys.path.insert(0, '.')
import __graft_entry__ as ge
ge.load_pkg()
from bwa_meme_b200 import synth, index_build, capi, prmi_train
out = []
for G in (20_000_000, 248_956_422):
    g = synth.make_genome_large(G, n_families=30, copies=100)
    text = index_build.fwd_rc_text(g)
    t0 = time.time(); r = capi.sa_build(text); t1 = time.time()
    bits = prmi_train.default_bits(len(text))
    l1, l2, st = capi.prmi_train(r["keys"], bits); t2 = time.time()
    kt = torch.from_numpy(r["keys"].view(np.int64)).cuda()
    torch.cuda.synchronize(); t3 = time.time()
    a1, a2, st2 = prmi_train.train(kt, bits); torch.cuda.synchronize(); t4 = time.time()
    pos, below, above = prmi_train.lookup_numpy(l1, l2, r["keys"][::97], len(text))
    d = pos - np.arange(len(text))[::97]
    ok = bool((d <= below).all() and (-d <= above).all())
    rec = {"suffixes": len(text), "bits": bits, "native_sa_keys_s": round(t1 - t0, 2), "sa_rounds": r["rounds"], "native_prmi_s": round(t2 - t1, 2),
           "torch_prmi_s": round(t4 - t3, 2), "native_stats": st, "torch_stats": {k: st2[k] for k in ("n_dense_leaf", "n_sub", "max_window")},
           "mean_window_native": float(((l2["err"] >> np.uint64(32)) + (l2["err"] & np.uint64(0xffffffff)))[(l2["err"] >> np.uint64(63)) == 0].mean()),
           "bounds_hold_on_sample": ok}
    print(json.dumps(rec), flush=True); out.append(rec)
    del kt; torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/r2_index_native.json", "w"), indent=1)
P
