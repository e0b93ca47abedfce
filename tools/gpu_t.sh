#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "invalid_utterance or non_finite" > $O/$1_pytest_edges.log 2>&1; tail -n 15 $O/$1_pytest_edges.log
