#!/bin/bash
# what the driver runs at round end, in its order: GPU tests in ONE process, smoke(), the reference arm, the bench arm
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/$1_pytest_all.log 2>&1; echo "rc=$?" >> $O/$1_pytest_all.log; tail -n 3 $O/$1_pytest_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 300 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 | cut -c1-300
timeout 400 python bench.py --gpus 1 --steps 50 --warmup 5 > $O/$1_bench.json 2> $O/$1_bench.err; python -c "
import json;b=json.load(open('$O/$1_bench.json'));print({k:b[k] for k in ('value','ms_per_step','gpu_launches','parity_checked','clocks')}, b['roofline']['frac'], b['roofline']['traffic_source'][:40], b['e2e']['value'], b['extra']['noise_scaled_mas']['ms_per_step'])"
