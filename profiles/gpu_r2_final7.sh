#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final7.log
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_bench_final7.json 2> gpurun_out/r2_bench_final7.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None, d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
print('mate', {k:v for k,v in d.get('mate_rescue',{}).items() if k in ('gcups_kernel','call_ms','kernel_ms','speedup_call_vs_cpu','parity_sample','error')})
a=d.get('align_e2e',{})
for k in ('se','pe'):
    if k in a: print(k, a[k]['reference_cpu'], a[k]['gpu']['Mbases_per_s'], a[k]['sam_identical'], a[k]['gpu'].get('chunk_s'))
if 'error' in a: print(a)
P
python -c "
import __graft_entry__ as g
g.smoke()" 2>&1 | tail -2
isa=avx512bw; grep -q " avx512bw" /proc/cpuinfo || isa=avx2
ls /root/.cache 2>/dev/null | head -2
FQ=$(ls /tmp/meme_b200_bench/align-*/se2000000.fq 2>/dev/null | head -1); echo "fq=$FQ"
if [ -n "$FQ" ]; then D=$(dirname $FQ); MEME_B200_LIB=$PWD/bwa-meme_b200/libmeme_b200.so MEME_B200=1 MEME_B200_CHAIN=1 MEME_B200_BSW=1 MEME_B200_SAM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cigar_align -s 2 -c 1 -f -o gpurun_out/ncu_k_cigar_align_batched oracle/_ref/bwa-meme-b200_mode3.$isa mem -7 -t 16 -K 100000000 $D/ref.fa $FQ -o /tmp/ncu.sam > gpurun_out/ncu_k_cigar_align_batched.log 2>&1; echo "ncu cigar rc=$?"; fi
