"""Start time of every stage for the three compute warps of one sub-partition (ktrace.py json): does a start skew persist?"""
import json, sys
d = json.load(open(sys.argv[1]))
ws = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 8, 12]
t = {w: {} for w in ws}
for tt, w, e, a in d["events"]:
    if w in t and e == 1:
        t[w][a] = tt
for n in sorted(t[ws[0]]):
    row = [t[w].get(n) for w in ws]
    print(n, row, [None if r is None or row[0] is None else r - row[0] for r in row])
