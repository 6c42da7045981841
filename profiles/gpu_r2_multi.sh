#!/bin/bash
# two GPUs: the drop-in path through the C ABI (SAM identical), end-to-end timing, and the torchrun bench with the record gather timed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_e2e_sam_gpu.py tests/test_seed_gpu.py::test_rmi_inference_matches_reference_formula tests/test_seed_gpu.py::test_probe_counts_stay_near_one_block_per_search -x -q 2>&1 | tail -4
ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_1gpu.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r2_e2e_se_2M_1gpu.jsonl
E2E_VERBOSE=1 DEVICES=0,1 ARMS=gpu_seed_bsw N=2000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_2gpu.jsonl 2> gpurun_out/r2_e2e_se_2M_2gpu.err; cut -c1-330 gpurun_out/r2_e2e_se_2M_2gpu.jsonl; grep "meme_b200" gpurun_out/r2_e2e_se_2M_2gpu.err | tail -4 | cut -c1-250
PE=1 DEVICES=0,1 ARMS=gpu_seed_bsw N=1000000 timeout 900 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_2gpu.jsonl 2>/dev/null; cut -c1-330 gpurun_out/r2_e2e_pe_1M_2gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err; cut -c1-900 gpurun_out/r2_bench_2gpu.json; python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1]); print('gather', json.dumps(j['results']['gather'])[:600]); print('e2e', j['e2e'])
except Exception as e: print('parse failed', e)
PY
