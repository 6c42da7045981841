#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2 3; do
for mode in on off; do
  if [ $mode = on ]; then export MEME_B200_OVERLAP=1; else unset MEME_B200_OVERLAP; fi
  E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_overlap2_pe_$mode.jsonl 2> gpurun_out/r2_overlap2_pe_$mode.err
  echo "PE overlap $mode: $(tail -1 gpurun_out/r2_overlap2_pe_$mode.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "global align" gpurun_out/r2_overlap2_pe_$mode.err | cut -c80-160 | tr '\n' ';'; echo
done; done
