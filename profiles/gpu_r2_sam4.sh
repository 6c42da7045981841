#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cigar_gpu.py -x -q 2>&1 | tail -4
MEME_CIGAR_DEBUG=1 E2E_VERBOSE=1 E2E_PROFILE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_sam4.jsonl 2> gpurun_out/r2_e2e_se_2M_sam4.err; echo "se rc=$?"; cut -c1-250 gpurun_out/r2_e2e_se_2M_sam4.jsonl; grep "meme_cigar\]\|global align\|SAM stage" gpurun_out/r2_e2e_se_2M_sam4.err | cut -c1-300
E2E_PROFILE=1 PE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_sam4.jsonl 2>gpurun_out/r2_e2e_pe_1M_sam4.err; echo "pe rc=$?"; cut -c1-250 gpurun_out/r2_e2e_pe_1M_sam4.jsonl; grep "SAM stage" gpurun_out/r2_e2e_pe_1M_sam4.err
