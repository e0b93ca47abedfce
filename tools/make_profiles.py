"""Turn the ncu artefacts a gpu_round.sh call left in gpurun_out/ into the tracked summaries under profiles/
(reads .ncu-rep files with `ncu -i`, no GPU needed)."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, SRC = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled"]

def summarise(rep, top=30):
    lines = []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    vals = {}
    for r in rows[2:]:
        lines.append("== kernel: " + r[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, r):
            if any(h == k or h.startswith(k) for k in KEYS):
                if "pcsamp" in h and (h.endswith("_not_issued") or v in ("0", "")):
                    continue
                lines.append(f"  {h} [{u}] = {v}")
                vals[h] = (v, u)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if hi:
        h = rows[hi[0]]
        data = [r for r in rows[hi[0] + 1:] if len(r) == len(h)]
        isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(int(r[ism]) for r in data if r[ism].isdigit())
        lines.append(f"== source: {len(data)} SASS lines, {tot} samples; hottest by samples:")
        for n, r in sorted(enumerate(data), key=lambda t: -int(t[1][ism] or 0))[:top]:
            lines.append(f"  line {n:5d} samples {r[ism]:>6} exec {r[iex]:>9}  {r[isrc][:100]}")
        ops = {}
        for r in data:
            op = r[isrc].split()[0] if r[isrc].split() else ""
            if op.startswith("@"):
                op = r[isrc].split()[1]
            op = op.split(".")[0]
            if op in ("UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "HMMA", "SYNCS"):
                ops[op] = ops.get(op, 0) + 1
        lines.append("== Blackwell-native SASS mnemonics present (static count): " + ", ".join(f"{k} x{v}" for k, v in sorted(ops.items())))
    return lines, vals

def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

for name in ("fused", "dp", "cost"):
    rep = os.path.join(SRC, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    lines, vals = summarise(rep)
    open(os.path.join(OUT, f"{tag}_{name}_ncu.txt"), "w").write("\n".join(lines) + "\n")
    if name == "fused" and "dram__bytes_read.sum" in vals:
        rd, wr = to_bytes(*vals["dram__bytes_read.sum"]), to_bytes(*vals["dram__bytes_write.sum"])
        json.dump({"kernel": "mas_fused_pair_kernel", "source": f"profiles/{tag}_fused_ncu.txt (ncu --set full, one launch, config 2)",
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr},
                  open(os.path.join(OUT, f"{tag}_fused_traffic.json"), "w"), indent=1)
    print("wrote", name)
for f in (f"{tag}_launches.csv", f"{tag}_bench.json", f"{tag}_bench_ref.json", f"{tag}_pytest.log", f"{tag}_smoke.log"):
    if os.path.exists(os.path.join(SRC, f)):
        shutil.copyfile(os.path.join(SRC, f), os.path.join(OUT, f))
if os.path.exists(os.path.join(OUT, f"{tag}_launches.csv")):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(OUT, f"{tag}_launches.csv")],
                       capture_output=True, text=True).stdout
    open(os.path.join(OUT, f"{tag}_launches_summary.txt"), "w").write(r)
    print(r)
