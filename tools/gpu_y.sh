#!/bin/bash
O=gpurun_out; mkdir -p $O
for f in test_gpu_round2 test_gpu_configs test_gpu_align; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x > $O/$1_pytest_$f.log 2>&1; tail -n 2 $O/$1_pytest_$f.log
done
timeout 600 python tools/bench_configs.py c1 c2 c3 c4 c5 --no-cpu --json $O/$1_configs.json > $O/$1_configs.log 2>&1; tail -n 6 $O/$1_configs.log
