#!/bin/bash
# regrouped seeding kernel on the small profiling workload: pass statistics, then one ncu --set full capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MEME_SEED_STATS=1 MEME_SEED_DEBUG=1 timeout 600 python profiles/prof_seed_small.py > gpurun_out/r2_small_stats.log 2>&1
grep "meme_seed\] T" gpurun_out/r2_small_stats.log | tail -3
MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grp -s 3 -c 3 -f -o gpurun_out/r2_prof_grp python profiles/prof_seed_small.py > gpurun_out/r2_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
