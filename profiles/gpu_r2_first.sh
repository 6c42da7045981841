#!/bin/bash
# round 2, first GPU contact of the regrouped seeding kernel: parity tests, then both kernels on the bench workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
for impl in grp flat; do
  MEME_SEED_IMPL=$impl MEME_SEED_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_$impl.json 2> gpurun_out/r2_bench_$impl.err
  echo "bench $impl rc=$?"; grep "meme_seed\] T" gpurun_out/r2_bench_$impl.err | tail -6; cat gpurun_out/r2_bench_$impl.json | cut -c1-300
done
for w in 8 16; do
  MEME_GRP_WARPS=$w MEME_SEED_DEBUG=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_grp_w$w.json 2> gpurun_out/r2_bench_grp_w$w.err
  echo "bench grp warps=$w rc=$?"; grep "meme_seed\] T" gpurun_out/r2_bench_grp_w$w.err | tail -3; cut -c1-200 gpurun_out/r2_bench_grp_w$w.json
done
