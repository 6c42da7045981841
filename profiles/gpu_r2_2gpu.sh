#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_e2e_sam_gpu.py -x -q -k "two_gpus" 2>&1 | tail -5
DEVICES=0,1 E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_2gpu_all.jsonl 2> gpurun_out/r2_e2e_se_2M_2gpu_all.err; echo "se rc=$?"; cut -c1-260 gpurun_out/r2_e2e_se_2M_2gpu_all.jsonl
DEVICES=0,1 E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_2gpu_all.jsonl 2> gpurun_out/r2_e2e_pe_1M_2gpu_all.err; echo "pe rc=$?"; cut -c1-260 gpurun_out/r2_e2e_pe_1M_2gpu_all.jsonl
grep "meme_b200\]" gpurun_out/r2_e2e_pe_1M_2gpu_all.err | tail -8 | cut -c1-250
