#!/bin/bash
# One GPU call of round 2: parity tests (new files first, each under its own timeout: a spin-wait bug must not take
# the whole call down), smoke, per-config numbers, bench (both arms), variants, compute-sanitizer.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round2.sh [tag]
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
for f in test_gpu_round2 test_gpu_configs test_gpu_align test_gpu_mas test_gpu_expand; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s -x > $O/${TAG}_pytest_$f.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_$f.log
  tail -4 $O/${TAG}_pytest_$f.log
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 900 python tools/bench_configs.py c1 c2 c3 c4 c5 --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; cat $O/${TAG}_configs.log | tail -8
MAS_DP_WARPS=4 timeout 300 python tools/bench_configs.py c1 c2 c3 --no-cpu --json $O/${TAG}_configs_w4.json > $O/${TAG}_configs_w4.log 2>&1; tail -4 $O/${TAG}_configs_w4.log
timeout 400 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cat $O/${TAG}_bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $O/${TAG}_sanitizer_$tool.log 2>&1; echo "rc=$?" >> $O/${TAG}_sanitizer_$tool.log
  tail -5 $O/${TAG}_sanitizer_$tool.log
done
