"""Summarise an .ncu-rep: headline metrics + hottest SASS lines (reads with `ncu -i`, no GPU needed)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__pcsamp_warps_issue_stalled"]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:90])
    for h, u, v in zip(hdr, units, r):
        if any(h == k or h.startswith(k) for k in keys):
            if "pcsamp" in h and (h.endswith("_not_issued") or v in ("0", "")):
                continue
            print(f"  {h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    hdr = rows[hi[0]]
    data = [r for r in rows[hi[0] + 1:] if len(r) == len(hdr)]
    isrc, iex, ism = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot_s = sum(int(r[ism]) for r in data if r[ism].isdigit())
    tot_e = sum(int(r[iex]) for r in data if r[iex].isdigit())
    print(f"== source: {len(data)} SASS lines, {tot_e} warp-instructions, {tot_s} samples; hottest by samples:")
    for n, r in sorted(enumerate(data), key=lambda t: -int(t[1][ism] or 0))[:top]:
        print(f"  line {n:5d} samples {r[ism]:>6} exec {r[iex]:>9}  {r[isrc][:90]}")
