"""Print the headline metrics of every kernel in an `ncu --page raw --csv` dump (developer tool)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "launch__registers_per_thread ", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum ", "launch__occupancy_limit",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max"]
for vals in rows[2:]:
    print("=====")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) or (w.endswith(" ") and h == w.strip()) for w in want):
            try:
                if "stalled" in h and float(v) < 0.3:
                    continue
            except ValueError:
                pass
            print(f"{h} [{u}] {v}")
