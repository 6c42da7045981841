#!/bin/bash
# SAM-stage global alignments on the GPU: parity (C ABI vs reference outputs vs oracle), e2e SAM, e2e timing with and without
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_cigar_gpu.py tests/test_e2e_sam_gpu.py -x -q > gpurun_out/r2_pytest_sam.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_sam.log
E2E_VERBOSE=1 E2E_PROFILE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_sam.jsonl 2> gpurun_out/r2_e2e_se_2M_sam.err; echo "se rc=$?"; cut -c1-330 gpurun_out/r2_e2e_se_2M_sam.jsonl; grep "meme_b200\] \(global\|ext\)" gpurun_out/r2_e2e_se_2M_sam.err | tail -6 | cut -c1-300; grep "WORKER_SAM" gpurun_out/r2_e2e_se_2M_sam.err
PE=1 ARMS=gpu_seed_chain_bsw,gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_sam.jsonl 2>gpurun_out/r2_e2e_pe_1M_sam.err; echo "pe rc=$?"; cut -c1-330 gpurun_out/r2_e2e_pe_1M_sam.jsonl
