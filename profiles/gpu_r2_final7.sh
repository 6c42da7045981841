#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final10.log
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_bench_final10.json 2> gpurun_out/r2_bench_final10.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final10.json').read().strip().splitlines()[-1])
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
