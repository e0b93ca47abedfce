#!/bin/bash
O=gpurun_out; mkdir -p $O
for c in c2 c3 c4; do for nw in 1 2 3; do for np in 1 2 4; do for st in 2 3 4; do
  [ $((nw*np)) -gt 8 ] && continue
  echo -n "$c nw=$nw np=$np st=$st "; MAS_SEG_NW=$nw MAS_SEG_PARTS=$np MAS_SEG_STAGES=$st timeout 120 python tools/bench_expand.py $c | grep -o '"expand_prior_backward_us[^,]*'
done; done; done; done 2>&1 | tee $O/$1_sweep.txt
