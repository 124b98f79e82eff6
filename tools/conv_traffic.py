"""Condenses the ncu metric pass over the conv fprop/dgrad launches of one C2 step into conv_traffic.json:
mean DRAM bytes (read + write) per launch, used by bench.py as roofline.traffic. Usage: conv_traffic.py in.csv out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, st = r, i + 1
        break
k, m, v, u, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
d = {}
for r in rows[st:]:
    if len(r) <= v:
        continue
    d.setdefault(r[idc], {"name": r[k].split("(")[0]})[r[m]] = float(r[v].replace(",", "")) * scale.get(r[u], 1.0)
n = len(d)
tot = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in d.values())
t = sum(x.get("gpu__time_duration.sum", 0) for x in d.values())
out = {"launches": n, "dram_bytes_per_launch": tot / max(n, 1), "dram_bytes_total": tot, "ncu_time_s_total": t,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the conv fprop+dgrad launches of one C2 training step"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
