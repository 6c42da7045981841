#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2_bench_final2.json 2> gpurun_out/r2_bench_final2.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None)
print('mate', {k:v for k,v in d.get('mate_rescue',{}).items() if k in ('gcups_kernel','call_ms','kernel_ms','speedup_call_vs_cpu','parity_sample','error')})
a=d.get('align_e2e',{})
for k in ('se','pe'):
    if k in a: print(k, a[k]['reference_cpu']['Mbases_per_s'], a[k]['gpu']['Mbases_per_s'], a[k]['sam_identical'])
if 'error' in a: print(a)
P
