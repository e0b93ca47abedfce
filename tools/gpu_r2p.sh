#!/bin/bash
TAG=${1:-r2p}
O=gpurun_out
mkdir -p $O
TRACELIB=$PWD/torch_tts_b200/libmas_b200_trace.so
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "noise" > $O/${TAG}_pytest_noise.log 2>&1; tail -n 5 $O/${TAG}_pytest_noise.log
timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -q -x > $O/${TAG}_pytest_configs.log 2>&1; tail -n 3 $O/${TAG}_pytest_configs.log
timeout 300 python tools/bench_configs.py c1 c2 --no-cpu --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -n 3 $O/${TAG}_configs.log
MAS_LIB_PATH=$TRACELIB timeout 150 python tools/trace_noise_fused.py > $O/${TAG}_trace_noise.txt 2>&1
grep -v "^cta" $O/${TAG}_trace_noise.txt | head -40
