#!/bin/bash
# A/B helper: each argument is a quoted list of VAR=value settings; runs the bench once per setting.
# Usage (GPU box): bash tools/ab.sh "MAS_DP_VK=0" "MAS_DP_VK=1 MAS_FUSED_ROUNDS=2"
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step']*1e3,2), {k:round(v*1e3,1) for k,v in d.get('kernels_ms',{}).items()})"
done
