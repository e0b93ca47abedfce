#!/bin/bash
# Full GPU validation of the current tree: every -m gpu test file (each under its own timeout), smoke, per-config
# numbers, bench (both arms).  Usage: bash tools/gpu_full.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
for f in test_gpu_round2 test_gpu_configs test_gpu_align test_gpu_mas test_gpu_expand; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -s -x > $O/${TAG}_pytest_$f.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_$f.log
  tail -n 3 $O/${TAG}_pytest_$f.log
done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -n 1 $O/${TAG}_smoke.log
for c in c1 c2 c3 c4 c5; do timeout 120 python tools/bench_expand.py $c; done > $O/${TAG}_expand.log 2>&1; tail -n 6 $O/${TAG}_expand.log
timeout 900 python tools/bench_configs.py c1 c2 c3 c4 c5 --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -n 8 $O/${TAG}_configs.log
timeout 400 python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; cut -c1-1200 $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cut -c1-600 $O/${TAG}_bench_ref.json
