#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_expand.py -m gpu -q -x > $O/$1_pytest_expand.log 2>&1; tail -n 3 $O/$1_pytest_expand.log
for c in c2 c3 c4; do timeout 120 python tools/bench_expand.py $c 2>&1 | tail -n 1; done; MAS_EXPAND_SCAN=0 timeout 120 python tools/bench_expand.py c2 2>&1 | tail -n 1
