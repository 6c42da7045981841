#!/bin/bash
# chaining on the GPU: parity (C ABI vs oracle vs reference dumps), e2e SAM + chain dumps, e2e timing with and without it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_chain_gpu.py tests/test_e2e_sam_gpu.py -x -q > gpurun_out/r2_pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_chain.log
E2E_VERBOSE=1 E2E_PROFILE=1 ARMS=gpu_seed_bsw,gpu_seed_chain_bsw N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_chain.jsonl 2> gpurun_out/r2_e2e_se_2M_chain.err; echo "se rc=$?"; cut -c1-400 gpurun_out/r2_e2e_se_2M_chain.jsonl; grep "meme_b200\] \(seeded\|ext\)" gpurun_out/r2_e2e_se_2M_chain.err | tail -6 | cut -c1-300
PE=1 ARMS=gpu_seed_bsw,gpu_seed_chain_bsw N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_chain.jsonl 2>gpurun_out/r2_e2e_pe_1M_chain.err; echo "pe rc=$?"; cut -c1-400 gpurun_out/r2_e2e_pe_1M_chain.jsonl
