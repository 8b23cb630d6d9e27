"""Digest of a ktrace.py timeline: where the compute warps and the producer of CTA 0 spend their cycles."""
import json, sys
from collections import defaultdict
import numpy as np

for path in sys.argv[1:]:
    d = json.load(open(path))
    per_warp = defaultdict(list)
    for t, w, e, a in d["events"]:
        if e != 63:
            per_warp[w].append((t, e, a))
    names = {(0, 1): "wait_full", (1, 2): "hdr", (2, 6): "fill", (6, 3): "sweep", (3, 2): "segswitch", (3, 4): "tail", (4, 5): "store",
             (5, 0): "loop", (1, 4): "inactive", (4, 0): "loop"}
    tot = defaultdict(int)
    fills, sw, switch, hdr = [], [], [], []
    for w, lst in per_warp.items():
        if w < 4:
            continue
        last = None
        nst = 0
        for (t, e, a) in lst:
            if last is not None:
                lt, le, la = last
                tot[names.get((le, e), "%d->%d" % (le, e))] += t - lt
                if (le, e) == (2, 6):
                    fills.append(t - lt); nst = la
                if (le, e) == (6, 3):
                    sw.append((t - lt, nst))
                if (le, e) == (3, 2):
                    switch.append(t - lt)
                if (le, e) == (1, 2):
                    hdr.append(t - lt)
            last = (t, e, a)
    T = sum(tot.values())
    span = max(t for t, _, _, _ in d["events"]) - min(t for t, _, _, _ in d["events"])
    print(d["workload"], "span", span, "compute warps:", {k: round(100 * v / T, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])})
    sw = np.array(sw)
    print("  per segment: fill %.0f  switch %.0f  hdr/stage %.0f  sweep %.1f cycles/step (%.1f steps/segment, %d segments/warp)" % (
        np.mean(fills), np.mean(switch) if switch else 0, np.mean(hdr), sw[:, 0].sum() / sw[:, 1].sum(), sw[:, 1].mean(), len(sw) // 8))
    lst = per_warp[0]
    cat = defaultdict(int)
    last = None
    pn = {(10, 11): "wait_landed", (11, 12): "patch", (12, 13): "to_wait_empty", (13, 14): "wait_empty", (14, 15): "issue", (15, 10): "next_stage",
          (12, 10): "next_stage(issued early)"}
    for (t, e, a) in lst:
        if last is not None:
            lt, le = last
            cat[pn.get((le, e), "%d->%d" % (le, e))] += t - lt
        last = (t, e)
    total = sum(cat.values())
    print("  producer:", {k: round(100 * v / total, 1) for k, v in sorted(cat.items(), key=lambda kv: -kv[1])})
    lst = per_warp[4]
    stages, cur = [], {}
    for (t, e, a) in lst:
        if e == 0:
            cur = {"t0": t}
        elif e == 1:
            cur["t1"] = t
        elif e == 4:
            cur["t4"] = t
        elif e == 5 and "t4" in cur:
            cur["t5"] = t
            stages.append(cur)
    print("  warp 4 stages (wait, compute, store):", [(s["t1"] - s["t0"], s["t4"] - s["t1"], s["t5"] - s["t4"]) for s in stages[:20]])
