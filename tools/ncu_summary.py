"""One-kernel summary (markdown) of an `ncu --set full` report: key raw metrics + hottest source lines.

    python tools/ncu_summary.py report.ncu-rep "title / command" > profiles/rNN_<kernel>_ncu.md
"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
print("# %s\n" % title)
for r in rows[2:]:
    print("## `%s`\n" % r[hdr.index("Kernel Name")][:100])
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            print("| %s | %s | %s |" % (k, r[hdr.index(k)], units[hdr.index(k)]))
    print()
src = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
print("## hottest source lines (share of executed warp instructions / of stall samples)\n\n```\n%s```" % src)
