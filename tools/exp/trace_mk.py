"""Digest of a ktrace.py timeline of the masked kernel with the TMEM store hand-off: producer warp 0 and compute warp 4."""
import json, sys
from collections import defaultdict
for path in sys.argv[1:]:
    d = json.load(open(path))
    per_warp = defaultdict(list)
    for t, w, e, a in d["events"]:
        if e != 63:
            per_warp[w].append((t, e, a))
    import os
    for w in [int(x) for x in os.environ.get("TRACE_WARPS", "0,4").split(",")]:
        cat = defaultdict(int)
        cnt = defaultdict(int)
        last = None
        for (t, e, a) in per_warp[w]:
            if last is not None:
                cat[(last[1], e)] += t - last[0]
                cnt[(last[1], e)] += 1
            last = (t, e)
        tot = sum(cat.values())
        print(d["workload"], "warp", w, "total", tot, {("%d->%d" % k): (round(100 * v / tot, 1), v // max(cnt[k], 1)) for k, v in sorted(cat.items(), key=lambda kv: -kv[1])})
