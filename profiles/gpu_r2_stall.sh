#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for rep in 1 2 3 4 5; do
  MEME_CIGAR_DEBUG=1 E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_stall_se_$rep.jsonl 2> gpurun_out/r2_stall_se_$rep.err
  echo "SE $rep: $(tail -1 gpurun_out/r2_stall_se_$rep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "meme_cigar\]\|global align" gpurun_out/r2_stall_se_$rep.err | head -5 | cut -c1-330
done
