#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final3.log
E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_pe_1M_final3.jsonl 2>gpurun_out/r2_e2e_pe_1M_final3.err; echo "pe rc=$?"; cut -c1-250 gpurun_out/r2_e2e_pe_1M_final3.jsonl; grep "mate rescue\|global align" gpurun_out/r2_e2e_pe_1M_final3.err | cut -c1-250
MEME_B200_NO_OVERLAP=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py 2>/dev/null | cut -c1-250
for c in 1048576 524288 262144; do
  MEME_SEED_CHUNK=$c timeout 900 python bench.py --steps 3 --warmup 3 --no-align --no-wide --no-chain --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk', $c, 'value', round(d['value']/1e6, 2), 'e2e', round(d['e2e']['value']/1e6, 2), 'ms', round(d['e2e']['ms_per_step'], 1))"
done
