for pair in 1 0; do for n in 64 48 32; do
echo "pair=$pair n_dp=$n: $(MAS_TC_PAIR=$pair MAS_FUSED_DP_CTAS=$n timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"]*1e3,1), "us/step", d["kernels_ms"])')"
done; done
