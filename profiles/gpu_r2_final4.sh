#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cigar_gpu.py tests/test_e2e_sam_gpu.py -x -q 2>&1 | tail -5
MEME_CIGAR_DEBUG=1 E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_final4.jsonl 2>gpurun_out/r2_e2e_se_2M_final4.err; echo "se rc=$?"; cut -c1-250 gpurun_out/r2_e2e_se_2M_final4.jsonl; grep "meme_cigar\]\|global align\|codes" gpurun_out/r2_e2e_se_2M_final4.err | cut -c1-300
