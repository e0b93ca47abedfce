#!/bin/bash
O=gpurun_out; mkdir -p $O
for n in 0 32 48 64 96; do
  for r in -1; do
  echo "== MAS_FUSED_DP_CTAS=$n"
  MAS_FUSED_DP_CTAS=$n timeout 300 python tools/bench_configs.py c3 c5 --no-cpu --json $O/$1_dpctas_$n.json 2>&1 | grep "^c[35]" | cut -c1-170
  done
done
