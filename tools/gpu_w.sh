#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_round2.py tests/test_gpu_mas.py -m gpu -q -x > $O/$1_pytest.log 2>&1; tail -n 5 $O/$1_pytest.log
timeout 600 python tools/bench_configs.py c4 --no-cpu --json $O/$1_configs.json > $O/$1_configs.log 2>&1; tail -n 3 $O/$1_configs.log
