"""One markdown table row per kernel launch in an .ncu-rep (developer tool): the metrics the design notes quote."""
import csv, subprocess, sys
COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__inst_executed.sum", "warp instr")]
STALL = "smsp__average_warps_issue_stalled_"
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"\n### {rep}\n")
    print("| kernel | " + " | ".join(n for _, n in COLS) + " | top stalls (warps per issue) |")
    print("|---|" + "---|" * (len(COLS) + 1))
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for m, _ in COLS:
            if m in ix:
                v, u = r[ix[m]], units[ix[m]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("-")
        st = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[ix[h]]), h[len(STALL):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st = ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:4] if n != "selected")
        print(f"| `{name}` | " + " | ".join(cells) + f" | {st} |")
