#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_round2.py -m gpu -q -x -s -k "config4 or wide_text or unaligned_rows" > $O/$1_pytest_c4.log 2>&1; grep -E "launches=|passed|failed|Error|assert" $O/$1_pytest_c4.log | head -8
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_align.py -m gpu -q -x > $O/$1_pytest_rest.log 2>&1; tail -n 3 $O/$1_pytest_rest.log
timeout 600 python tools/bench_configs.py c1 c2 c3 c4 c5 --no-cpu --json $O/$1_configs.json > $O/$1_configs.log 2>&1; tail -n 3 $O/$1_configs.log
