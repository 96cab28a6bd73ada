"""print the per-warp event trace written by VADC_BWD_TRACE for tiles [a, b): python scripts/trace_view.py file a b"""
import sys, collections
ev = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
a, b = int(sys.argv[2]), int(sys.argv[3])
t0 = min(e[3] for e in ev)
names = {0: "P unit top", 3: "P slot read + next load issued", 1: "P converted", 2: "P arrive", 10: "E1 top", 11: "E1 AREMPTY ok", 12: "E1 AFULL", 13: "E1 PFULL ok", 14: "E1 G1FULL ok",
         15: "E1 RFULL", 16: "E1 end", 19: "E3 top", 20: "E3 RFULL ok", 21: "E3 ACCFULL ok", 22: "E3 end", 40: "E1 G1 loaded", 41: "E1 computed", 42: "E1 REMPTY ok", 43: "E1 r stored", 44: "E1 fenced", 33: "MMA S5b", 30: "MMA S1", 31: "MMA S5a", 32: "MMA S3"}
for w, code, it, c in sorted(ev, key=lambda e: e[3]):
    if a <= it < b and (w in (0, 2, 14, 4, 15) or len(sys.argv) > 4):
        print(f"{c - t0:9d}  w{w:<2d} tile {it:3d}  {names.get(code, code)}")
# per-tile period
last = {}
for w, code, it, c in ev:
    if w == 0 and code == 15: last[it] = c
its = sorted(last)
print("E1 RFULL period (cycles):", [last[i + 1] - last[i] for i in its[:-1]][:30])
