"""kernel share of a step from an ncu launch list: python scripts/launch_shares.py launches.csv "<header comment>" """
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; kn = h.index("Kernel Name"); v = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    t = float(r[v].replace(",", "")) / 1e3
    a = agg.setdefault(r[kn], [0.0, 0])
    a[0] += t; a[1] += 1
tot = sum(a[0] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# total {tot / 1e3:.3f} ms over {sum(a[1] for a in agg.values())} launches")
print("share%  us/launch  launches  kernel")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * t / tot:6.2f}  {t / n:9.1f}  {n:8d}  {k[:110]}")
