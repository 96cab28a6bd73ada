"""sample histogram over SASS line buckets + stall reasons: python scripts/ncu_buckets.py rep [bucket]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; bk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, rr in enumerate(rows) if "# Samples" in rr)
h = rows[hi]; body = rows[hi + 1:]
ix = {n: i for i, n in enumerate(h)}
def iv(x):
    try: return int(float(x))
    except Exception: return 0
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(iv(b[ix["# Samples"]]) for b in body)
for s in range(0, len(body), bk):
    seg = body[s:s + bk]
    n = sum(iv(b[ix["# Samples"]]) for b in seg)
    ex = sum(iv(b[ix["Instructions Executed"]]) for b in seg)
    dom = sorted(((sum(iv(b[ix[k]]) for b in seg), k[6:]) for k in stalls), reverse=True)[:3]
    marks = [b[ix["Source"]].split()[0] for b in seg if any(t in b[ix["Source"]] for t in ("SYNCS", "BAR.", "UTCHMMA", "LDTM", "UTMASTG", "EXIT"))]
    print(f"{s:5d} {100*n/tot:5.1f}% exec={ex:>10d} {dom} {sorted(set(marks))}")
