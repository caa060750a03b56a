"""Sum dram__bytes_read/write per kernel from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv` log: one line per kernel name (launches, total MB, total us).  Usage: ncu_traffic.py log.csv [launches_per_step]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
acc = OrderedDict()
for r in rows[1:]:
    name = r[ci["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "")
    m, v, u = r[ci["Metric Name"]], float(r[ci["Metric Value"]].replace(",", "")), r[ci["Metric Unit"]]
    d = acc.setdefault(name, {"n": 0, "bytes": 0.0, "us": 0.0})
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if m.startswith("dram__bytes"):
        d["bytes"] += v * scale
    elif m.startswith("gpu__time"):
        d["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        d["n"] += 1
tot_b = tot_t = 0.0
for k, d in acc.items():
    print(f"{k:60s} x{d['n']:4d}  {d['bytes'] / 1e6:10.1f} MB  {d['us']:10.1f} us")
    tot_b += d["bytes"]
    tot_t += d["us"]
print(f"{'TOTAL':60s}        {tot_b / 1e6:10.1f} MB  {tot_t:10.1f} us")
