#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_final_2gpu.json 2> gpurun_out/r2_bench_final_2gpu.err; echo "bench2 rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_final_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['results']['gather']['record_gather'])
P
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200; echo "ref wall ${SECONDS}s"
