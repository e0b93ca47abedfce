#!/bin/bash
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
for f in test_gpu_round2 test_gpu_configs test_gpu_align; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x > $O/${TAG}_pytest_$f.log 2>&1; echo "rc=$?" >> $O/${TAG}_pytest_$f.log
  tail -n 3 $O/${TAG}_pytest_$f.log
done
timeout 600 python tools/bench_configs.py c1 c2 c3 c5 --no-cpu --json $O/${TAG}_configs.json > $O/${TAG}_configs.log 2>&1; tail -n 6 $O/${TAG}_configs.log
export MAS_LIB_PATH=$PWD/torch_tts_b200/libmas_b200_trace.so
timeout 150 python tools/trace_fused.py > $O/${TAG}_trace_fused.txt 2>&1
head -n 24 $O/${TAG}_trace_fused.txt; grep "^cta 70" $O/${TAG}_trace_fused.txt
