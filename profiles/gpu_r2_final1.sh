#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu_mate.log
N=400000 timeout 600 python profiles/prof_matesw.py > gpurun_out/r2_matesw_bench.json 2> gpurun_out/r2_matesw_bench.err; echo "matesw bench rc=$?"; cut -c1-1200 gpurun_out/r2_matesw_bench.json
E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_mate2.jsonl 2>gpurun_out/r2_e2e_pe_1M_mate2.err; echo "pe rc=$?"; cut -c1-250 gpurun_out/r2_e2e_pe_1M_mate2.jsonl; grep "mate rescue\|Processed\|reserved" gpurun_out/r2_e2e_pe_1M_mate2.err | cut -c1-250
E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_mate2.jsonl 2>gpurun_out/r2_e2e_se_2M_mate2.err; echo "se rc=$?"; cut -c1-250 gpurun_out/r2_e2e_se_2M_mate2.jsonl
# ncu: one launch each of the two new kernels
N=131072 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_matesw -c 1 -f -o gpurun_out/ncu_k_matesw python profiles/prof_matesw.py > gpurun_out/ncu_k_matesw.log 2>&1; echo "ncu matesw rc=$?"
isa=avx512bw; grep -q " avx512bw" /proc/cpuinfo || isa=avx2
MEME_B200_LIB=$PWD/bwa-meme_b200/libmeme_b200.so MEME_B200=1 MEME_B200_CHAIN=1 MEME_B200_BSW=1 MEME_B200_SAM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cigar_align -s 2 -c 1 -f -o gpurun_out/ncu_k_cigar_align oracle/_ref/bwa-meme-b200_mode3.$isa mem -7 -t 16 -K 100000000 /tmp/meme_e2e/ref.fa /tmp/meme_e2e/se2000000.fq -o /tmp/meme_e2e/ncu.sam > gpurun_out/ncu_k_cigar_align.log 2>&1; echo "ncu cigar rc=$?"
ls -la gpurun_out/*.ncu-rep
