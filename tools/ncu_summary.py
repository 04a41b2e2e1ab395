"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into profiles/:
   python tools/ncu_summary.py gpurun_out/launches.csv profiles/r01_launches  [step_index]
writes <out>_step.csv (every launch of one step: id, kernel, grid, block, us) and <out>_summary.md
(per-kernel totals and shares of that step)."""
import csv, sys

src, out = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -2
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    rows.append((int(x["ID"]), x["Kernel Name"].split("(")[0].replace("<unnamed>::", ""), x["Grid Size"], x["Block Size"], v))
starts = [i for i, r in enumerate(rows) if "pack_input" in r[1]]
if len(starts) < 2:
    raise SystemExit("need at least two steps in the launch list")
# a step = [noise draw, counter add, ... whatever precedes pack_input] + pack_input ... last optimizer launch:
# it begins right after the previous step's last optimizer kernel
def step_begin(i):
    j = i - 1
    while j >= 0 and "optim" not in rows[j][1]:
        j -= 1
    return j + 1 if j >= 0 else max(0, i - 4)
s = step_begin(starts[which])
e = step_begin(starts[which + 1]) if which + 1 != 0 and which + 1 < len(starts) else len(rows)
step = rows[s:e]
with open(out + "_step.csv", "w") as f:
    f.write("id,kernel,grid,block,us\n")
    for r in step:
        f.write(f'{r[0]},"{r[1]}","{r[2]}","{r[3]}",{r[4]:.3f}\n')
agg, tot = {}, 0.0
for r in step:
    a = agg.setdefault(r[1], [0.0, 0])
    a[0] += r[4]
    a[1] += 1
    tot += r[4]
with open(out + "_summary.md", "w") as f:
    f.write(f"# ncu launch list, one training step ({len(step)} launches, {tot:.1f} us serialised, cold cache)\n\n")
    f.write(f"source: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --no-graph)\n\n")
    f.write("| kernel | launches | us | share |\n|---|---|---|---|\n")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"| `{k[-70:]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
print(open(out + "_summary.md").read())
