#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_sam_gpu.py -x -q 2>&1 | tail -4
for rep in 1 2; do
for mode in 2 1; do
  export MEME_B200_EXT_HALVES=$((mode-1))
  E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_halves_se_$mode.jsonl 2> gpurun_out/r2_halves_se_$mode.err
  echo "SE pieces $mode: $(tail -1 gpurun_out/r2_halves_se_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "extended" gpurun_out/r2_halves_se_$mode.err | cut -c40-260
  E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_halves_pe_$mode.jsonl 2> gpurun_out/r2_halves_pe_$mode.err
  echo "PE pieces $mode: $(tail -1 gpurun_out/r2_halves_pe_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"
done; done
