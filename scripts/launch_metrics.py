"""per-launch time and DRAM bytes from an ncu --csv launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum; argument 2 = keep the last N launches"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
d = collections.OrderedDict()
for x in csv.DictReader(lines):
    d.setdefault((x["ID"], x["Kernel Name"][:90]), {})[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
items = list(d.items())
if len(sys.argv) > 2:
    items = items[-int(sys.argv[2]):]
tot = 0.0
for (i, k), m in items:
    t = m.get("gpu__time_duration.sum", 0) / 1e3
    tot += t
    print(f"{t:8.1f} us  rd {m.get('dram__bytes_read.sum', 0) / 1e6:7.1f} MB  wr {m.get('dram__bytes_write.sum', 0) / 1e6:7.1f} MB  {k}")
print(f"{tot:8.1f} us total over {len(items)} launches")
