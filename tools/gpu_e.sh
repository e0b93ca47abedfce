#!/bin/bash
# prior-expansion consumers: tests + device times at configs 1-5 (usage: bash tools/gpu_e.sh <tag>)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_expand.py -m gpu -q -x > $O/$1_pytest_expand.log 2>&1; tail -n 3 $O/$1_pytest_expand.log
for c in c1 c2 c3 c4 c5; do timeout 300 python tools/bench_expand.py $c; done 2>&1 | tee $O/$1_expand.jsonl
for st in 2 4; do for c in c2 c4; do echo -n "stages=$st "; MAS_SEG_STAGES=$st timeout 300 python tools/bench_expand.py $c; done; done 2>&1 | tee -a $O/$1_expand.jsonl
