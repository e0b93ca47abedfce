set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1
python bench.py --steps 50 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/r1_ncu_bench.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_smoke.log | tail -2; cat gpurun_out/r1_bench.json; cat gpurun_out/r1_bench_ref.json
