"""per-tile milestones of the backward's event trace: python scripts/trace_summary.py file [first] [count]"""
import sys, collections
ev = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
a = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
t0 = min(e[3] for e in ev)
by = collections.defaultdict(list)
for w, code, it, c in ev:
    by[(code, it)].append((c - t0, w))
def first(code, it): return min((c for c, _ in by.get((code, it), [(None, 0)]) if c is not None), default=None)
def last(code, it): return max((c for c, _ in by.get((code, it), [(None, 0)]) if c is not None), default=None)
cols = [("G first top", lambda t: first(0, t)), ("G last arrive", lambda t: last(2, t)), ("S1", lambda t: first(30, t)), ("S5a", lambda t: first(31, t)),
        ("E1 G1ok", lambda t: first(14, t)), ("E1 RFULL", lambda t: last(15, t)), ("S3", lambda t: first(32, t)),
        ("E3 ACCok", lambda t: first(21, t)), ("E3 end", lambda t: last(22, t)), ("S5b", lambda t: first(33, t))]
print("tile " + " ".join(f"{c[0]:>13s}" for c in cols))
prev = None
for t in range(a, a + n):
    vals = [f(t) for _, f in cols]
    print(f"{t:4d} " + " ".join(f"{(v if v is not None else -1):13d}" for v in vals))
print("deltas relative to S1 of the tile, and tile period (S1 to S1):")
for t in range(a, a + n):
    s1 = first(30, t)
    if s1 is None: continue
    nx = first(30, t + 1)
    d = {c[0]: (c[1](t) - s1 if c[1](t) is not None else None) for c in cols}
    print(t, d, "period", None if nx is None else nx - s1)
# busy estimate of E3: ACCok -> end per warp
for w in (4, 5, 8, 9, 12, 13):
    tops = {it: c for ww, code, it, c in ev if ww == w and code == 21}
    ends = {it: c for ww, code, it, c in ev if ww == w and code == 22}
    rf = {it: c for ww, code, it, c in ev if ww == w and code == 19}
    its = sorted(set(tops) & set(ends))
    if len(its) > 2:
        busy = sum(ends[i] - tops[i] for i in its)
        span = ends[its[-1]] - tops[its[0]]
        print(f"E3 warp {w}: busy {100 * busy / span:.0f}% of its span, mean {busy / len(its):.0f} cycles per tile, mean period {span / (len(its) - 1):.0f}")
for w in (2, 3):
    tops = {it: c for ww, code, it, c in ev if ww == w and code == 14}
    ends = {it: c for ww, code, it, c in ev if ww == w and code == 15}
    its = sorted(set(tops) & set(ends))
    if len(its) > 2:
        busy = sum(ends[i] - tops[i] for i in its); span = ends[its[-1]] - tops[its[0]]
        print(f"E1 warp {w}: G1ok->RFULL busy {100 * busy / span:.0f}%, mean {busy / len(its):.0f} cycles")
for w in (0, 1, 6, 7, 10, 11, 14):
    tops = [c for ww, code, it, c in ev if ww == w and code == 0]
    rd = [c for ww, code, it, c in ev if ww == w and code == 3]
    cv = [c for ww, code, it, c in ev if ww == w and code == 1]
    ar = [c for ww, code, it, c in ev if ww == w and code == 2]
    m = min(len(tops), len(rd), len(cv), len(ar))
    if m > 2:
        span = ar[m - 1] - tops[0]
        print(f"P warp {w}: units {m}, slot wait+read {sum(rd[i]-tops[i] for i in range(m))/m:.0f}, Gempty wait+convert {sum(cv[i]-rd[i] for i in range(m))/m:.0f}, fence+arrive {sum(ar[i]-cv[i] for i in range(m))/m:.0f}, per unit {span/m:.0f}")
