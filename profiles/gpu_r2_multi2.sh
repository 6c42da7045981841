#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu4.log
E2E_VERBOSE=1 DEVICES=0,1 ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_2gpu.jsonl 2> gpurun_out/r2_e2e_se_2M_2gpu.err; cut -c1-330 gpurun_out/r2_e2e_se_2M_2gpu.jsonl; grep "meme_b200" gpurun_out/r2_e2e_se_2M_2gpu.err | tail -3 | cut -c1-250
PE=1 DEVICES=0,1 ARMS=gpu_seed_bsw N=1000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_2gpu.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r2_e2e_pe_1M_2gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r2_bench_2gpu.err; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1]); print('value', j['value'], 'ms', j['ms_per_step']); print('gather', json.dumps(j['results']['gather']['record_gather'])[:400]); print('e2e', j['e2e'])
except Exception as e: print('parse failed', e)
PY
