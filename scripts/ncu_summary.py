"""Key metrics of `ncu --set full` captures (exported with `ncu -i X.ncu-rep --page raw --csv`) -> markdown table.
usage: python scripts/ncu_summary.py raw1.csv [raw2.csv ...]"""
import csv, sys

KEYS = [
    ("time us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic"),
    ("DRAM read MB", "dram__bytes_read.sum"),
    ("DRAM write MB", "dram__bytes_write.sum"),
    ("DRAM % of peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("smem bank conflicts (st)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("smem bank conflicts (ld)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("stall long_scoreboard /issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall barrier /issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall short_scoreboard /issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall wait /issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
]
cols = []
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        cols.append((name, d, dict(zip(hdr, units))))
print("| metric | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---:|" * len(cols))
for label, key in KEYS:
    vals = []
    for _, d, u in cols:
        v = d.get(key, "")
        try:
            f = float(v.replace(",", ""))
            if u.get(key) in ("Mbyte", "Kbyte", "Gbyte", "byte") and label.endswith("MB"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u[key]]
            v = f"{f:,.1f}" if abs(f) < 1e6 else f"{f:,.0f}"
        except ValueError:
            pass
        vals.append(v)
    print(f"| {label} | " + " | ".join(vals) + " |")
