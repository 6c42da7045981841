#!/bin/bash
# flat kernel with aligned 16-byte staging: parity, timing on the small workload (narrow and wide indexes), bench value, DRAM traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seed_gpu.py tests/test_text_end_guard.py -x -q -m gpu 2>&1 | tail -3
echo "== small workload, narrow"; MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3 | cut -c1-60
echo "== small workload, wide MINB=2"; MEME_FORCE_WIDE=1 MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3 | cut -c1-60
echo "== small workload, wide MINB=1"; MEME_SEED_WIDE_MINB=1 MEME_FORCE_WIDE=1 MEME_SEED_DEBUG=1 timeout 300 python profiles/prof_seed_small.py 2>&1 | grep "meme_seed\] T" | tail -3 | cut -c1-60
timeout 900 python bench.py --steps 5 --warmup 3 --no-align > gpurun_out/r2_bench_flat2.json 2> gpurun_out/r2_bench_flat2.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_flat2.json').read().strip().splitlines()[-1])
print('value', j['value'], 'ms', j['ms_per_step'], 'e2e', j['e2e']['value'], 'kernel_ms', j['kernel_ms'], 'wide', j['wide'], 'parity', j.get('parity_sample',{}).get('identical_to_oracle'))
PY
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:k_flat -c 3 --csv --log-file gpurun_out/r2_ncu_flat2_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-wide --no-align > /dev/null 2> gpurun_out/r2_ncu_flat2.err; echo "ncu rc=$?"; grep -E "k_flat" gpurun_out/r2_ncu_flat2_traffic.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-200 | head -20
