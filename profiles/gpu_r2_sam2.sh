#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MEME_CIGAR_DEBUG=1 E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_sam2.jsonl 2> gpurun_out/r2_e2e_se_2M_sam2.err; echo "se rc=$?"; cut -c1-250 gpurun_out/r2_e2e_se_2M_sam2.jsonl; grep "meme_cigar\]\|global align\|SAM stage" gpurun_out/r2_e2e_se_2M_sam2.err | cut -c1-300
