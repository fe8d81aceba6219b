"""Summarise an ncu --set full report (one kernel launch) into a small JSON for profiles/ (run here, no GPU needed):
   python scripts/ncu_summary.py gpurun_out/prof_k1.ncu-rep profiles/r2_ms_kernel_ncu_full.json "<note>" """
import csv, json, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "lts__t_sectors.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
d = {"report": rep, "note": note, "how": "ncu --set full --clock-control none --import-source on, one launch, cold cache"}
for h, u, v in zip(hdr, units, vals):
    if h in want:
        try:
            d[h + (" [" + u + "]" if u else "")] = float(v)
        except ValueError:
            d[h] = v
json.dump(d, open(out, "w"), indent=1)
print(out, d.get("gpu__time_duration.sum [us]"))
