#!/usr/bin/env python3
"""one line per kernel of an ncu_summary.py text file: python profiles/ncu_brief.py file.txt ..."""
import sys
COLS = [("t", "gpu__time_duration.sum"), ("rd", "dram__bytes_read.sum"), ("wr", "dram__bytes_write.sum"), ("regs", "launch__registers_per_thread"),
        ("warps", "sm__warps_active.avg.per_cycle_active"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        ("l1hit", "l1tex__t_sector_hit_rate.pct"), ("l2hit", "lts__t_sector_hit_rate.pct"),
        ("math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
        ("wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        ("long", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("short", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        ("mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
        ("lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio")]
for f in sys.argv[1:]:
    print(f)
    for b in open(f).read().split('---')[1:]:
        d = {}
        for line in b.strip().splitlines():
            parts = line.split()
            if len(parts) >= 2:
                d[parts[0]] = parts[1:]
        name = ' '.join(d.get('Kernel', ['?']))[5:40]
        def g(k):
            v = d.get(k, ['-'])
            try:
                return "%.3g" % float(v[0].replace(',', '')) + (v[1][0] if len(v) > 1 and v[1] in ("ms", "us", "Gbyte", "Mbyte") else "")
            except ValueError:
                return v[0]
        print("  %-36s" % name + " ".join("%s=%s" % (a, g(k)) for a, k in COLS))
