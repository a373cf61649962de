"""Key metrics of `ncu --set full` captures (exported with `ncu -i X.ncu-rep --page raw --csv`) -> markdown.

usage: python scripts/ncu_summary.py [--distinct] [--rows] [--trim out.csv] raw1.csv [raw2.csv ...]
  --distinct   one capture per kernel name: its longest launch
  --rows       one kernel per ROW (readable with many kernels); default: one kernel per column
  --trim F     also write the columns bench.py / this script read to F (small enough to commit under profiles/)
"""
import csv
import sys

KEYS = [
    ("time us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem KB", "launch__shared_mem_per_block_dynamic"),
    ("DRAM read MB", "dram__bytes_read.sum"),
    ("DRAM write MB", "dram__bytes_write.sum"),
    ("DRAM GB/s", "dram__bytes.sum.per_second"),
    ("DRAM % of peak", "+dram__bytes_read.sum.pct_of_peak_sustained_elapsed+dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warp instr (M)", "smsp__inst_executed.sum"),
    ("smem st conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
    ("smem ld conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
    ("stall long_sb /issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall barrier /issue", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall short_sb /issue", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
]
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3,
         "byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}


def value(d, u, label, key):
    try:
        if key.startswith("+"):
            f = sum(float(d[k].replace(",", "")) for k in key[1:].split("+"))
        else:
            f = float(d[key].replace(",", ""))
            if label.endswith("MB") or label.endswith("GB/s"):
                f *= SCALE.get(u.get(key), 1.0)
            if label == "dyn smem KB":
                f *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(u.get(key), 1.0)
            if label.endswith("(M)"):
                f *= 1e-6
    except (KeyError, ValueError):
        return ""
    return f"{f:,.1f}" if abs(f) < 1e6 else f"{f:,.0f}"


def main():
    args = sys.argv[1:]
    distinct, by_rows = "--distinct" in args, "--rows" in args
    trim = args[args.index("--trim") + 1] if "--trim" in args else None
    paths = [a for i, a in enumerate(args) if not a.startswith("--") and (i == 0 or args[i - 1] != "--trim")]
    cols, seen = [], set()
    needed = {"ID", "Kernel Name"}
    for _, k in KEYS:
        needed.update(k[1:].split("+") if k.startswith("+") else [k])
    trimmed = None
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        keep = [i for i, h in enumerate(hdr) if h in needed]
        if trim and trimmed is None:
            trimmed = [[hdr[i] for i in keep], [units[i] for i in keep]]
        body = [r for r in rows[2:] if len(r) >= len(hdr)]
        if distinct:
            # one capture per kernel name: the LONGEST launch (the step's main instance of that instantiation, not the
            # small head / embedding calls that happen to come first), in order of first appearance
            def dur(r):
                try:
                    return float(dict(zip(hdr, r))["gpu__time_duration.sum"].replace(",", ""))
                except (KeyError, ValueError):
                    return 0.0
            best, order = {}, []
            for r in body:
                name = dict(zip(hdr, r))["Kernel Name"].split("(")[0].replace("void ", "")
                if name not in best:
                    order.append(name)
                if name not in best or dur(r) > dur(best[name]):
                    best[name] = r
            body = sorted((best[n] for n in order), key=dur, reverse=True)
        for r in body:
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            cols.append((name, d, dict(zip(hdr, units))))
            if trimmed is not None:
                trimmed.append([r[i] for i in keep])
    if trim and trimmed:
        csv.writer(open(trim, "w", newline="")).writerows(trimmed)
    if by_rows:
        print("| kernel | " + " | ".join(l for l, _ in KEYS) + " |")
        print("|---|" + "---:|" * len(KEYS))
        for name, d, u in cols:
            print(f"| `{name}` | " + " | ".join(value(d, u, l, k) for l, k in KEYS) + " |")
    else:
        print("| metric | " + " | ".join(c[0] for c in cols) + " |")
        print("|---|" + "---:|" * len(cols))
        for label, key in KEYS:
            print(f"| {label} | " + " | ".join(value(d, u, label, key) for _, d, u in cols) + " |")


if __name__ == "__main__":
    main()
