#!/bin/bash
TAG=${1:-r2r}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "noise" > $O/${TAG}_pytest_noise.log 2>&1; tail -n 3 $O/${TAG}_pytest_noise.log
timeout 300 python tools/bench_configs.py c1 c2 --no-cpu --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -n 3 $O/${TAG}_configs.log
