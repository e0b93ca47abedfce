#!/bin/bash
O=gpurun_out; mkdir -p $O
MAS_LIB_PATH=torch_tts_b200/libmas_b200_trace.so timeout 300 python tools/debug_plain.py 2>&1 | grep "run \|deterministic" 
MAS_LIB_PATH=torch_tts_b200/libmas_b200_trace.so DB=160 DT=300 timeout 300 python tools/debug_plain.py 2>&1 | grep "run \|deterministic" 
timeout 600 python tools/bench_configs.py c1 c2 c3 c5 --no-cpu --json $O/$1_configs.json > $O/$1_configs.log 2>&1; tail -n 4 $O/$1_configs.log
