#!/bin/bash
TAG=${1:-r2f}
O=gpurun_out
mkdir -p $O
export MAS_LIB_PATH=$PWD/torch_tts_b200/libmas_b200_trace.so
for dbg in 0 64 128 192; do
MAS_DP_DEBUG=$dbg timeout 150 python tools/trace_noise_fused.py > $O/${TAG}_trace_noise_$dbg.txt 2>&1
echo "== debug $dbg"; grep -A40 "value warp 0" $O/${TAG}_trace_noise_$dbg.txt
done
