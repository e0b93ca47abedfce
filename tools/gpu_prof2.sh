#!/bin/bash
# Round-2 ncu evidence (one GPU): launch list of the bench arm, full captures of the fused kernel (no noise) and of the
# single-launch noise kernel, each after its plain run has exited 0.
O=gpurun_out; mkdir -p $O
BENCH="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline"
timeout 200 $BENCH > $O/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $O/r2_launches.csv $BENCH > $O/r2_ncu_launches.log 2>&1
timeout 200 $BENCH > $O/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_fused_pair -s 4 -c 1 -f -o $O/r2_fused $BENCH > $O/r2_ncu_fused.log 2>&1
python tools/run_once.py c2 noise > $O/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_fused_noise -s 3 -c 1 -f -o $O/r2_noise python tools/run_once.py c2 noise > $O/r2_ncu_noise.log 2>&1
python tools/run_once.py c2 noise > $O/r2_plain4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2_noise_launches.csv python tools/run_once.py c2 noise > $O/r2_ncu_noise_launches.log 2>&1
ls -la $O/r2_*.ncu-rep $O/r2_launches.csv $O/r2_noise_launches.csv
# the wide-text fused kernel (config 4) and the prior-expansion backward (config 2)
python tools/run_once.py c4 > $O/r2_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_fused_pair -s 3 -c 1 -f -o $O/r2_wide python tools/run_once.py c4 > $O/r2_ncu_wide.log 2>&1
python tools/bench_expand.py c2 > $O/r2_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mas_segsum -s 2 -c 1 -f -o $O/r2_segsum python tools/bench_expand.py c2 > $O/r2_ncu_segsum.log 2>&1
ls -la $O/r2_*.ncu-rep
