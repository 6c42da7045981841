#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MEME_CIGAR_DEBUG=1 E2E_VERBOSE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_final5.jsonl 2>gpurun_out/r2_e2e_se_2M_final5.err; echo "se rc=$?"; cut -c1-250 gpurun_out/r2_e2e_se_2M_final5.jsonl; grep "meme_cigar\]" gpurun_out/r2_e2e_se_2M_final5.err | cut -c1-300
E2E_VERBOSE=1 PE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_final5.jsonl 2>gpurun_out/r2_e2e_pe_1M_final5.err; echo "pe rc=$?"; cut -c1-250 gpurun_out/r2_e2e_pe_1M_final5.jsonl; grep "global align\|mate rescue" gpurun_out/r2_e2e_pe_1M_final5.err | cut -c1-300
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -3
