#!/bin/bash
TAG=${1:-r2c}
O=gpurun_out
mkdir -p $O
for f in test_gpu_round2 test_gpu_configs; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -s -x > $O/${TAG}_pytest_$f.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_$f.log
  tail -5 $O/${TAG}_pytest_$f.log
done
timeout 600 python tools/bench_configs.py c1 c2 c3 c5 --no-cpu --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -6 $O/${TAG}_configs.log
for k in 1 2 3; do
  MAS_FUSED_ROUNDS=$k timeout 300 python tools/bench_configs.py c2 --no-cpu --json $O/${TAG}_configs_k$k.json > $O/${TAG}_configs_k$k.log 2>&1; echo "rounds=$k"; tail -1 $O/${TAG}_configs_k$k.log
done
export MAS_LIB_PATH=$PWD/torch_tts_b200/libmas_b200_trace.so
timeout 150 python tools/trace_noise_fused.py > $O/${TAG}_trace_noise.txt 2>&1
timeout 150 python tools/trace_fused.py > $O/${TAG}_trace_fused.txt 2>&1
cat $O/${TAG}_trace_noise.txt
