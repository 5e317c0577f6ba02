"""Turn an .ncu-rep (one kernel, --set full --import-source on) into a compact text summary for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_tma_ld.sum", "sm__cycles_elapsed.max"]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name)
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"  {h:78s} {units[i]:14s} {vals[i]}")
    print("  -- warp stall (issue-stalled per issue-active, > 0.2)")
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith("ratio"):
            try:
                v = float(vals[i])
            except ValueError:
                continue
            if v > 0.2:
                print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = {}
    for r in data:
        for k in hdr:
            if k.startswith("stall_") and "Not" not in k and r[ix[k]] not in ("", "0"):
                agg[k] = agg.get(k, 0) + int(r[ix[k]])
    print(f"  -- pc sampling: {tot} samples;", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print("  -- hottest instructions")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:8]:
        print(f"  {int(r[ix['# Samples']]):7d}  {r[ix['Source']].strip()[:80]}")
