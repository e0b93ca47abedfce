#!/bin/bash
# usage: bash tools/gpu_ncu.sh <tag> <kernel-regex> <run_once args...>
TAG=$1; PAT=$2; shift 2
O=gpurun_out; mkdir -p $O
python tools/run_once.py "$@" > $O/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$PAT -s 3 -c 1 -f -o $O/${TAG} python tools/run_once.py "$@" > $O/${TAG}_ncu.log 2>&1
tail -n 3 $O/${TAG}_plain.log; tail -n 3 $O/${TAG}_ncu.log
