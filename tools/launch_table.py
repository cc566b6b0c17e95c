"""Print an ncu --csv launch list (gpu__time_duration + dram bytes) as one line per launch.
usage: python tools/launch_table.py file.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i
        break
d = OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    rec = dict(zip(hdr, r))
    d.setdefault((int(rec["ID"]), rec["Kernel Name"]), {})[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
tot = 0.0
for (i, name), m in d.items():
    if i < skip:
        continue
    short = name.split("(")[0].replace("void ", "")[:90]
    t = m.get("gpu__time_duration.sum", 0) / 1e6
    tot += t
    print(f"{i:4d} {t:8.3f} ms  rd {m.get('dram__bytes_read.sum', 0)/1e9:7.3f} GB  wr {m.get('dram__bytes_write.sum', 0)/1e9:7.3f} GB  {short}")
print(f"total {tot:.3f} ms")
