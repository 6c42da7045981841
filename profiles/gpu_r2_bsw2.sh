#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_gpu3.log
N=2000000 DUMP=/tmp/meme_e2e/pairs timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M.jsonl 2> gpurun_out/r2_e2e_se_2M.err; echo "e2e SE rc=$?"; cut -c1-700 gpurun_out/r2_e2e_se_2M.jsonl; tail -3 gpurun_out/r2_e2e_se_2M.err
timeout 300 python profiles/bench_bsw.py /tmp/meme_e2e/pairs.left /tmp/meme_e2e/pairs.right > gpurun_out/r2_bsw_real_pairs.jsonl 2> gpurun_out/r2_bsw_real_pairs.err; echo "bsw bench rc=$?"; cat gpurun_out/r2_bsw_real_pairs.jsonl; tail -3 gpurun_out/r2_bsw_real_pairs.err
for side in left right; do timeout 300 oracle/_ref/bsw_shim_mode3.avx512bw /tmp/meme_e2e/pairs.$side - bench $(nproc) 2 2>&1 | tail -1; done | tee gpurun_out/r2_bsw_cpu_baseline.jsonl
PE=1 N=1000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M.jsonl 2> gpurun_out/r2_e2e_pe_1M.err; echo "e2e PE rc=$?"; cut -c1-700 gpurun_out/r2_e2e_pe_1M.jsonl; tail -3 gpurun_out/r2_e2e_pe_1M.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bsw_is -c 6 -f -o gpurun_out/r2_prof_bsw_is python profiles/bench_bsw.py /tmp/meme_e2e/pairs.right > gpurun_out/r2_ncu_bsw.log 2>&1; echo "ncu rc=$?"
