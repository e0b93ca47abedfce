#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python tools/bench_configs.py c2 --no-cpu --json $O/$1_configs.json > $O/$1_configs.log 2>&1; tail -n 2 $O/$1_configs.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/$1_bench.json 2> $O/$1_bench.err; python -c "
import json;b=json.load(open('$O/$1_bench.json'));print(b['ms_per_step'],b['kernels_ms'],b['extra']['noise_scaled_mas']['ms_per_step'],b['parity_checked'])"
