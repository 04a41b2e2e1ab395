"""DRAM traffic of the tensor-core GEMM family over one training step, from an ncu metrics pass:
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
       --log-file gpurun_out/dram.csv python bench.py --no-graph --no-cpu --no-gpu-eager --no-sustained --steps 2 --warmup 3
   python tools/dram_family.py gpurun_out/dram.csv profiles/gemm_family_dram.json [step_index] [bench config key]"""
import csv, json, sys
src, out = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lines = [l for l in open(src) if not l.startswith("==")]
per = {}
order = []
for x in csv.DictReader(lines):
    i = int(x["ID"])
    if i not in per:
        per[i] = {"name": x["Kernel Name"].split("(")[0].replace("<unnamed>::", "")}
        order.append(i)
    v = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    m = x["Metric Name"]
    if m == "gpu__time_duration.sum":
        per[i]["us"] = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        per[i][m] = v * mult
rows = [per[i] for i in order]
starts = [k for k, r in enumerate(rows) if "pack_input" in r["name"]]
s, e = starts[which], starts[which + 1]
fam = [r for r in rows[s:e] if r["name"].startswith(("gemm_tc", "wgrad_tc"))]
rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in fam)
wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in fam)
doc = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, "
                 "python bench.py --no-graph --no-cpu --no-gpu-eager --no-sustained --steps 2 --warmup 3 (one eager step, "
                 "cold-cache serialised launches)",
       "kernels": "gemm_tc_kernel + gemm_tc_bnr_kernel + wgrad_tc_kernel", "launches": len(fam),
       "dram_bytes_read": rd, "dram_bytes_write": wr, "device_us": sum(r["us"] for r in fam),
       "dram_bytes_per_step": rd + wr,
       "per_kernel": {n: {"launches": sum(1 for r in fam if r["name"] == n),
                          "dram_bytes": sum(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0) for r in fam if r["name"] == n),
                          "us": sum(r["us"] for r in fam if r["name"] == n)} for n in sorted({r["name"] for r in fam})}}
cfg = sys.argv[4] if len(sys.argv) > 4 else "2"
try:
    full = json.load(open(out))
except Exception:
    full = {}
full[cfg] = doc  # bench.py reads profiles/gemm_family_dram.json[<--config>]["dram_bytes_per_step"]
json.dump(full, open(out, "w"), indent=1)
print(json.dumps(doc, indent=1))
