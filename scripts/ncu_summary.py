"""Print the key metrics of every kernel in an .ncu-rep (reads `ncu -i <rep> --page raw --csv`)."""
import csv, subprocess, sys
out = open(sys.argv[1]).read() if sys.argv[1].endswith(".csv") else \
    subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum"]
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    if name in seen and "--all" not in sys.argv:
        continue
    seen.add(name)
    print("====", name, "grid", r[idx.get("Grid Size", 0)], "block", r[idx.get("Block Size", 0)])
    for k in KEYS:
        if k in idx:
            print(f"  {k:72s} {r[idx[k]]:>16s} {units[idx[k]]}")
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h:
            try:
                stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
    print("  top stalls:", ", ".join(f"{n.split('stalled_')[-1]} {v:.0f}%" for v, n in sorted(stalls, reverse=True)[:5]))
