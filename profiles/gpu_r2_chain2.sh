#!/bin/bash
# chaining in the default GPU arm: whole GPU suite, full bench line (chain record, align_e2e with chaining), launch list of the timed kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_chain.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_chain.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1800 python bench.py > gpurun_out/r2_bench_chain_1gpu.json 2> gpurun_out/r2_bench_chain_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_chain_1gpu.err | cut -c1-200
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_chain_1gpu.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'frac', j['roofline']['frac'], 'cpu', j['cpu_baseline']['value'], 'wide', j['wide']['value'])
print('chain', j['chain'])
a=j['align_e2e']; print('align se', a['se']['reference_cpu']['Mbases_per_s'], a['se']['gpu']['Mbases_per_s'], a['se']['sam_identical'], 'pe', a['pe']['reference_cpu']['Mbases_per_s'], a['pe']['gpu']['Mbases_per_s'], a['pe']['sam_identical'])
print(a['se']['gpu']['log_tail'])
b=j['bsw']; print('bsw gcups', b['gcups'], 'pairs/s', b['pairs_per_s'])
PY
E2E_VERBOSE=1 ARMS=gpu_seed_chain_bsw N=2000000 timeout 1200 python profiles/e2e_align.py > gpurun_out/r2_e2e_se_2M_chain2.jsonl 2> gpurun_out/r2_e2e_se_2M_chain2.err; grep "meme_b200\] \(seeded\|ext\|warm\)" gpurun_out/r2_e2e_se_2M_chain2.err | tail -8 | cut -c1-300; cut -c1-300 gpurun_out/r2_e2e_se_2M_chain2.jsonl
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_flat|k_pack|k_gather|k_counts|k_chain|DeviceRadixSort|DeviceScan' -c 300 --csv --log-file gpurun_out/r2_ncu_launches_bench2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-align --no-wide > /dev/null 2> gpurun_out/r2_ncu_launches_bench2.err; echo "ncu rc=$?"; wc -l gpurun_out/r2_ncu_launches_bench2.csv
