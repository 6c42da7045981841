#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matesw_gpu.py tests/test_cigar_gpu.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_e2e_sam_gpu.py -x -q -k "pe or paired" 2>&1 | tail -15
E2E_VERBOSE=1 E2E_PROFILE=1 PE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_mate1.jsonl 2>gpurun_out/r2_e2e_pe_1M_mate1.err; echo "pe rc=$?"; cut -c1-250 gpurun_out/r2_e2e_pe_1M_mate1.jsonl; grep "mate rescue\|Processed" gpurun_out/r2_e2e_pe_1M_mate1.err | cut -c1-250
