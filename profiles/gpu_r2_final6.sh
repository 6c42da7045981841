#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final6.log
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_bench_final6.json 2> gpurun_out/r2_bench_final6.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None, d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
print('mate', {k:v for k,v in d.get('mate_rescue',{}).items() if k in ('gcups_kernel','call_ms','kernel_ms','speedup_call_vs_cpu','parity_sample','error')})
a=d.get('align_e2e',{})
for k in ('se','pe'):
    if k in a: print(k, a[k]['reference_cpu']['Mbases_per_s'], a[k]['gpu']['Mbases_per_s'], a[k]['sam_identical'], a[k]['gpu'].get('chunk_s'), a[k]['reference_cpu'].get('chunk_s'))
if 'error' in a: print(a)
print('bsw', d.get('bsw',{}).get('gcups'), 'wide', d.get('wide',{}).get('value'), 'chain', d.get('chain',{}).get('k_chain_ms'))
P
grep -i "index built\|genome" gpurun_out/r2_bench_final6.err | head -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_final6.csv python bench.py --steps 2 --warmup 1 --no-align --no-cpu-baseline --no-wide > gpurun_out/r2_launches_bench_final6.out 2>/dev/null; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_launches_bench_final6.csv
