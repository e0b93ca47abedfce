#!/bin/bash
# multi-GPU evidence: bash tools/gpu_multi.sh <N> <tag>
N=$1; TAG=${2:-r2}
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $RUN tools/run_sharded_check.py > $O/${TAG}_sharded_check_n$N.log 2>&1; tail -n 2 $O/${TAG}_sharded_check_n$N.log
timeout 300 $RUN tools/bench_c5_sharded.py --json $O/${TAG}_c5_sharded_n$N.json > $O/${TAG}_c5_sharded_n$N.log 2>&1; tail -n 1 $O/${TAG}_c5_sharded_n$N.log | cut -c1-700
timeout 300 $RUN tools/bench_c5_sharded.py --noise --json $O/${TAG}_c5_sharded_noise_n$N.json > $O/${TAG}_c5_sharded_noise_n$N.log 2>&1; tail -n 1 $O/${TAG}_c5_sharded_noise_n$N.log | cut -c1-700
timeout 300 $RUN bench.py --gpus $N --steps 50 --warmup 5 > $O/${TAG}_scale$N.json 2> $O/${TAG}_scale$N.err; cut -c1-400 $O/${TAG}_scale$N.json
