#!/bin/bash
# round-2 validation on one B200: all GPU tests, smoke, the full bench line, the reference arm, the ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_final_1gpu.err | cut -c1-200
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_final_1gpu.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'cpu', j['cpu_baseline']['value'], 'wide', j['wide']['value'])
a=j['align_e2e']; print('align se', a['se']['reference_cpu']['Mbases_per_s'], a['se']['gpu']['Mbases_per_s'], a['se']['sam_identical'], 'pe', a['pe']['reference_cpu']['Mbases_per_s'], a['pe']['gpu']['Mbases_per_s'], a['pe']['sam_identical'])
b=j['bsw']; print('bsw gcups', b['gcups'], 'pairs/s', b['pairs_per_s'], 'cpu', b.get('cpu_baseline'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_bench_final_reference.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-align > /dev/null 2> gpurun_out/r2_ncu_launches_bench.err; echo "ncu rc=$?"; wc -l gpurun_out/r2_ncu_launches_bench.csv
