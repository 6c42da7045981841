#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cigar_gpu.py tests/test_matesw_gpu.py tests/test_e2e_sam_gpu.py -x -q 2>&1 | tail -4
for rep in 1 2 3; do
  E2E_VERBOSE=1 ARMS=gpu_all N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_final8_se_$rep.jsonl 2> gpurun_out/r2_final8_se_$rep.err
  echo "SE $rep: $(tail -1 gpurun_out/r2_final8_se_$rep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "global align" gpurun_out/r2_final8_se_$rep.err | cut -c60-140 | tr '\n' ';'; echo
  E2E_VERBOSE=1 PE=1 ARMS=gpu_all N=1000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_final8_pe_$rep.jsonl 2> gpurun_out/r2_final8_pe_$rep.err
  echo "PE $rep: $(tail -1 gpurun_out/r2_final8_pe_$rep.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['Mbases_per_s'], d['sam_identical_to_reference'])")"; grep "global align\|Smith-Waterman pairs" gpurun_out/r2_final8_pe_$rep.err | cut -c60-160 | tr '\n' ';'; echo
done
