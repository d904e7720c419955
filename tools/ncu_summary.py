#!/usr/bin/env python
"""One line of key metrics per kernel from `ncu -i x.ncu-rep --page raw --csv` (the files kept under profiles/)."""
import csv, sys
KEYS = [("gpu__time_duration.sum", "dur"), ("smsp__inst_executed.sum", "winst"), ("smsp__issue_active.avg.per_cycle_active", "issue"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1wave%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); un = dict(zip(h, u))
        print("%s  [%s]" % (d.get("Kernel Name", "?")[:90], path.split("/")[-1]))
        print("   " + "  ".join("%s=%s%s" % (n, d[k][:9], {"Mbyte": "MB", "Gbyte": "GB", "Kbyte": "KB", "us": "us", "ms": "ms"}.get(un.get(k, ""), "")) for k, n in KEYS if k in d))
