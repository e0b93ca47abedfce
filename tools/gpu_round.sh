#!/bin/bash
# One GPU call: parity tests, smoke, bench (both arms), ncu launch list and full captures of the top kernels.
# Usage (from the repo root, on the GPU box): bash tools/gpu_round.sh
mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
BENCH="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline"
timeout 120 $BENCH > gpurun_out/r1_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1_launches.csv $BENCH > gpurun_out/r1_ncu_launches.log 2>&1
timeout 120 $BENCH > gpurun_out/r1_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_fused -s 4 -c 1 -f -o gpurun_out/r1_fused $BENCH > gpurun_out/r1_ncu_fused.log 2>&1
timeout 120 $BENCH > gpurun_out/r1_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_dp_kernel -s 4 -c 1 -f -o gpurun_out/r1_dp $BENCH > gpurun_out/r1_ncu_dp.log 2>&1
timeout 120 $BENCH > gpurun_out/r1_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_cost_tc -s 4 -c 1 -f -o gpurun_out/r1_cost $BENCH > gpurun_out/r1_ncu_cost.log 2>&1
set +x
tail -3 gpurun_out/r1_pytest.log; tail -2 gpurun_out/r1_smoke.log; cat gpurun_out/r1_bench.json; cat gpurun_out/r1_bench_ref.json
