#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
E2E_VERBOSE=1 ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_v2.jsonl 2> gpurun_out/r2_e2e_se_2M_v2.err; echo "e2e SE rc=$?"; cut -c1-300 gpurun_out/r2_e2e_se_2M_v2.jsonl; grep "meme_b200" gpurun_out/r2_e2e_se_2M_v2.err | cut -c1-260
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_full.err; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_full.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','wide','align_e2e','bsw','kernel_ms'):
        print(k, json.dumps(j.get(k))[:1200])
except Exception as e: print('parse failed', e)
PY
