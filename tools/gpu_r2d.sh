#!/bin/bash
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k noise > $O/${TAG}_pytest_noise.log 2>&1; tail -3 $O/${TAG}_pytest_noise.log
timeout 600 python tools/bench_configs.py c1 c2 c3 c5 --no-cpu --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -6 $O/${TAG}_configs.log
export MAS_LIB_PATH=$PWD/torch_tts_b200/libmas_b200_trace.so
timeout 150 python tools/trace_noise_fused.py > $O/${TAG}_trace_noise.txt 2>&1
MAS_FUSED_ROUNDS=2 timeout 150 python tools/trace_fused.py > $O/${TAG}_trace_fused_k2.txt 2>&1
cat $O/${TAG}_trace_noise.txt
