#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PE=1 ARMS=none N=1000000 timeout 1200 python profiles/e2e_align.py > /dev/null 2>&1     # builds the index and the FASTQs, runs the reference arm only
isa=avx512bw; grep -q " avx512bw" /proc/cpuinfo || isa=avx2
MEME_B200_LIB=$PWD/bwa-meme_b200/libmeme_b200.so MEME_B200=1 MEME_B200_CHAIN=1 MEME_B200_BSW=1 MEME_B200_SAM=1 MEME_B200_MATE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_aligner_pe.csv oracle/_ref/bwa-meme-b200_mode3.$isa mem -7 -t 16 -K 100000000 /tmp/meme_e2e/ref.fa /tmp/meme_e2e/pe1000000_1.fq /tmp/meme_e2e/pe1000000_2.fq -o /tmp/meme_e2e/ncu_pe.sam > gpurun_out/r2_launches_aligner_pe.out 2>&1; echo "rc=$?"; wc -l gpurun_out/r2_launches_aligner_pe.csv; grep "Processed" gpurun_out/r2_launches_aligner_pe.out | tail -3
