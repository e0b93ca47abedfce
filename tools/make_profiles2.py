"""Round 2: turn the artefacts the GPU calls left in gpurun_out/ into the tracked summaries under profiles/
(reads .ncu-rep files with `ncu -i`, no GPU needed)."""
import importlib.util, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, OUT = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.argv = [sys.argv[0], "__none__"]          # make_profiles.py copies nothing for an unknown tag
spec = importlib.util.spec_from_file_location("make_profiles", os.path.join(ROOT, "tools", "make_profiles.py"))
mp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mp)
KERNELS = {"fused": ("mas_fused_pair_kernel", "config 2"), "noise": ("mas_fused_noise_kernel", "config 2, noise-scaled"),
           "wide": ("mas_fused_pair_kernel (W = 4 teams, 3 column blocks)", "config 4"),
           "segsum": ("mas_segsum_kernel", "config 2, prior-expansion backward")}
for name in KERNELS:
    rep = f"{SRC}/r2_{name}.ncu-rep"
    if not os.path.exists(rep):
        continue
    lines, vals = mp.summarise(rep)
    open(f"{OUT}/r2_{name}_ncu.txt", "w").write("\n".join(lines) + "\n")
    rd, wr = mp.to_bytes(*vals["dram__bytes_read.sum"]), mp.to_bytes(*vals["dram__bytes_write.sum"])
    kern, what = KERNELS[name]
    json.dump({"kernel": kern, "source": f"profiles/r2_{name}_ncu.txt (ncu --set full --clock-control none, one launch, {what})",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr},
              open(f"{OUT}/r2_{name}_traffic.json", "w"), indent=1)
    print(name, "dram MB read/write", rd / 1e6, wr / 1e6)
cp = {
    "r2f_bench.json": "r2_bench.json", "r2f_bench_ref.json": "r2_bench_ref.json", "r2f_configs.json": "r2_configs.json",
    "r2f_configs.log": "r2_configs.log", "r2f_expand.log": "r2_expand.log", "r2f_smoke.log": "r2_smoke.log",
    "r2_launches.csv": "r2_launches.csv", "r2_noise_launches.csv": "r2_noise_launches.csv",
    "r2_scale2.json": "r2_scale2.json", "r2_scale4.json": "r2_scale4.json", "r2_scale8.json": "r2_scale8.json",
    "r2_sharded_check_n2.log": "r2_sharded_check_n2.log", "r2_sharded_check_n4.log": "r2_sharded_check_n4.log",
    "r2_sharded_check_n8.log": "r2_sharded_check_n8.log",
    "r2t_trace_noise.txt": "r2_trace_noise.txt", "r2h_trace_fused.txt": "r2_trace_fused.txt",
    "r2a_sanitizer_memcheck.log": "r2_sanitizer_closed.log",
}
for n in (2, 4, 8):
    cp[f"r2_c5_sharded_n{n}.json"] = f"r2_c5_sharded_n{n}.json"
    cp[f"r2_c5_sharded_noise_n{n}.json"] = f"r2_c5_sharded_noise_n{n}.json"
for a, b in cp.items():
    if os.path.exists(f"{SRC}/{a}"):
        shutil.copyfile(f"{SRC}/{a}", f"{OUT}/{b}")
    else:
        print("missing", a)
with open(f"{OUT}/r2_pytest.log", "w") as f:
    for n in ("round2", "configs", "align", "mas", "expand"):
        p = f"{SRC}/r2f_pytest_test_gpu_{n}.log"
        if os.path.exists(p):
            f.write(f"== tests/test_gpu_{n}.py\n" + open(p).read())
for n in ("r2_launches", "r2_noise_launches"):
    if os.path.exists(f"{OUT}/{n}.csv"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), f"{OUT}/{n}.csv"],
                           capture_output=True, text=True).stdout
        open(f"{OUT}/{n}_summary.txt", "w").write(r)
