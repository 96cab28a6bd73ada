"""print the top stall locations of an ncu source-page CSV (ncu -i rep --page source --csv)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
body = rows[2:]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for i in sorted(top):
    r = body[i]
    s = int(r[ix["# Samples"]])
    dom = sorted(((int(r[ix[n]]), n) for n in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*s/tot:5.1f}% exec={r[ix['Instructions Executed']]:>8} {r[ix['Source']].strip()[:90]:90s} {dom}")
