"""One line per launch of an .ncu-rep: python tools/ncu_compact.py rep.ncu-rep "title" > profiles/x.md"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, u = r[0], r[1]
def col(name):
    return h.index(name) if name in h else None
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp insts")]
idx = [(col(c), n, c) for c, n in cols if col(c) is not None]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else rep}\n")
print("| " + " | ".join(n + (f" [{u[i]}]" if u[i] else "") for i, n, _ in idx) + " |")
print("|" + "---|" * len(idx))
for row in r[2:]:
    vals = []
    for i, n, c in idx:
        v = row[i]
        if c == "Kernel Name":
            v = v.split("(")[0].replace("<unnamed>::", "")
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        vals.append(v)
    print("| " + " | ".join(vals) + " |")
