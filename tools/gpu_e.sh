#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_expand.py -m gpu -q -x > $O/$1_pytest_expand.log 2>&1; tail -n 12 $O/$1_pytest_expand.log
for c in c1 c2 c3 c4 c5; do timeout 300 python tools/bench_expand.py $c; done 2>&1 | tee $O/$1_expand.jsonl
for c in c1 c2 c4; do MAS_SEG_PARTS=2 timeout 300 python tools/bench_expand.py $c; done 2>&1 | tee -a $O/$1_expand.jsonl
