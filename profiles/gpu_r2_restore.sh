#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
MEME_B200_MATE_RESTORE=1 timeout 900 python -m pytest tests/test_e2e_sam_gpu.py -x -q 2>&1 | tail -4
for rep in 1 2; do
for mode in once twice; do
  if [ $mode = once ]; then export MEME_B200_MATE_RESTORE=1; else unset MEME_B200_MATE_RESTORE; fi
  E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_restore_pe_$mode.jsonl 2> gpurun_out/r2_restore_pe_$mode.err
  echo "PE pre $mode: $(tail -1 gpurun_out/r2_restore_pe_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"
done; done
grep "mate rescue" gpurun_out/r2_restore_pe_once.err | tail -2
