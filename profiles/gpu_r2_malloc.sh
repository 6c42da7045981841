#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2; do
for mode in on off; do
  if [ $mode = off ]; then export MEME_B200_NO_MALLOPT=1; else unset MEME_B200_NO_MALLOPT; fi
  E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_malloc_se_$mode.jsonl 2> gpurun_out/r2_malloc_se_$mode.err
  echo "SE mallopt $mode: $(tail -1 gpurun_out/r2_malloc_se_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "extended" gpurun_out/r2_malloc_se_$mode.err | cut -c60-200
  E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_malloc_pe_$mode.jsonl 2> gpurun_out/r2_malloc_pe_$mode.err
  echo "PE mallopt $mode: $(tail -1 gpurun_out/r2_malloc_pe_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"
done; done
