#!/bin/bash
O=gpurun_out; mkdir -p $O
MAS_LIB_PATH=$PWD/torch_tts_b200/libmas_b200_trace.so timeout 150 python tools/trace_noise_fused.py > $O/$1_trace_noise.txt 2>&1
grep -v "^cta" $O/$1_trace_noise.txt | tail -32
