timeout 120 python tools/check_cost.py | grep -v "us/call"; echo rc=$?
timeout 100 python tools/time_cost.py 0 16 18 20
MAS_TC_PAIR=0 timeout 100 python tools/time_cost.py 16
MAS_TC_GRID=84 timeout 60 python tools/time_cost.py 16
timeout 200 python -m pytest tests/test_gpu_align.py -x -q 2>&1 | tail -5
bash tools/_sweep.sh
