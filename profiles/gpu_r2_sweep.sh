#!/bin/bash
# regrouped kernel v2 (converged interpreter, four pools): parity, then a (warps, threshold) sweep on the small workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seed_gpu.py -x -q 2>&1 | tail -3
MEME_SEED_IMPL=flat MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3 | cut -c1-60
for w in 8 12 16 24 32; do for th in 8 24 48 96; do
  echo "== warps $w th $th"
  MEME_GRP_WARPS=$w MEME_GRP_TH=$th MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3 | cut -c1-60
done; done
MEME_SEED_STATS=1 MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grp -s 3 -c 3 -f -o gpurun_out/r2_prof_grp2 python profiles/prof_seed_small.py > gpurun_out/r2_ncu2.log 2>&1
echo "ncu rc=$?"
