#!/bin/bash
# kernel B with the reference side read from the resident text: parity, e2e SAM + timing, aligner profile table, configs[0]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bsw_gpu.py tests/test_e2e_sam_gpu.py -x -q 2>&1 | tail -4
E2E_VERBOSE=1 E2E_PROFILE=1 ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_text.jsonl 2> gpurun_out/r2_e2e_se_2M_text.err; cut -c1-330 gpurun_out/r2_e2e_se_2M_text.jsonl; grep "meme_b200\] ext" gpurun_out/r2_e2e_se_2M_text.err | tail -3 | cut -c1-250
MEME_B200_BSW_BYTES=1 ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py 2>/dev/null | tail -1 | cut -c1-330
PE=1 ARMS=gpu_seed_bsw N=1000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_text.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r2_e2e_pe_1M_text.jsonl
rm -rf /tmp/meme_e2e; MODE=1 G=4600000 N=100000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_config0_mode1.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r2_e2e_config0_mode1.jsonl
