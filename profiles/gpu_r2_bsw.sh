#!/bin/bash
# round 2: inter-sequence BSW kernel + chunk-level extension driver: parity tests, end-to-end SAM diff + timing, kernel-B bench on real pairs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu2.log
N=1000000 DUMP=/tmp/meme_e2e/pairs timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_1M.jsonl 2> gpurun_out/r2_e2e_se_1M.err; echo "e2e SE rc=$?"; cut -c1-420 gpurun_out/r2_e2e_se_1M.jsonl; tail -3 gpurun_out/r2_e2e_se_1M.err
ls -la /tmp/meme_e2e/pairs* 2>/dev/null
timeout 300 python profiles/bench_bsw.py /tmp/meme_e2e/pairs.left /tmp/meme_e2e/pairs.right > gpurun_out/r2_bsw_real_pairs.jsonl 2> gpurun_out/r2_bsw_real_pairs.err; echo "bsw bench rc=$?"; cat gpurun_out/r2_bsw_real_pairs.jsonl; tail -3 gpurun_out/r2_bsw_real_pairs.err
PE=1 N=500000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_500k.jsonl 2> gpurun_out/r2_e2e_pe_500k.err; echo "e2e PE rc=$?"; cut -c1-420 gpurun_out/r2_e2e_pe_500k.jsonl; tail -3 gpurun_out/r2_e2e_pe_500k.err
