"""CPU model of the tiled kernel's tap program (round-2 design study; mirrors the device builder in csrc/taps.cu).

Dense sheared column groups: the PSF support is sheared by k columns per row (x' = x - k (y - ymin)), cut into groups of
G sheared columns, groups into bands, bands into row windows (chunks).  A segment = one group's rows [first, last] inside
a chunk; every step of a segment executes G taps densely (zero weights where the PSF has none).  Reports fill efficiency
(useful taps / executed slots) and register reuse (FMAs per shared-memory load)."""
import sys
import numpy as np

C = 7


def band_groups(G, k, R, chunk_rows, pitch_slack=64, cap_cols=24):
    cols = min(cap_cols, pitch_slack + 1 - (R - 1) * abs(k) - (chunk_rows - 1) * abs(k))
    return max(1, cols // G)


def plan(psf, G, k, R=8, chunk_rows=20):
    ys, xs = np.nonzero(psf)
    ymin = ys.min()
    xp = xs - k * (ys - ymin)
    x0 = xp.min()
    g = (xp - x0) // G
    ngroups = g.max() + 1
    nb = band_groups(G, k, R, chunk_rows)
    steps = segs = nch = 0
    for b0 in range(0, ngroups, nb):
        inband = (g >= b0) & (g < b0 + nb)
        if not inband.any():
            continue
        rows_band = np.unique(ys[inband])
        cursor = rows_band.min()
        while True:
            rem = rows_band[rows_band >= cursor]
            if len(rem) == 0:
                break
            y0 = rem.min()
            y1 = y0 + chunk_rows - 1
            nch += 1
            for gg in range(b0, min(b0 + nb, ngroups)):
                sel = (g == gg) & (ys >= y0) & (ys <= y1)
                if sel.any():
                    segs += 1
                    steps += ys[sel].max() - ys[sel].min() + 1
            cursor = y1 + 1
    taps = len(ys)
    wl = C + G - 1
    loads = segs * (R - 1) * wl + steps * (wl + 1)
    ffma2 = steps * G * (R // 2) * C
    other = steps * (wl + 6) + segs * ((R - 1) * wl + 30) + nch * 40
    return dict(taps=taps, steps=steps, segs=segs, chunks=nch, loads=loads, ffma2=ffma2, other=other,
                eff=taps / (steps * G), reuse=taps * R * C / loads, G=G, k=k,
                cost=max(2 * ffma2, ffma2 + other))


def best(psf, R=8, kmax=1, Gs=(2, 4)):
    cands = [plan(psf, G, k, R) for G in Gs for k in range(-kmax, kmax + 1)]
    return min(cands, key=lambda p: p["cost"])


if __name__ == "__main__":
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "/tmp/psfs/psfs.npz")
    for R, kmax in ((8, 1), (8, 0), (6, 2), (6, 0)):
        for name in ("c2", "c3"):
            tot = dict(taps=0, steps=0, segs=0, loads=0, ffma2=0, other=0, cost=0, slots=0, chunks=0)
            desc = []
            for psf in d[name]:
                b = best(psf, R, kmax)
                b["slots"] = b["steps"] * b["G"]
                desc.append("G%dk%+d" % (b["G"], b["k"]))
                for key in tot:
                    tot[key] += b[key]
            ideal = tot["taps"] * (R // 2) * C * 2
            print("R%d kmax%d %s: eff %.3f reuse %.2f  fma-cycles/ideal %.2f  cost/ideal %.2f smem@fma-bound %.0f%% chunks %d | %s" % (
                R, kmax, name, tot["taps"] / tot["slots"], tot["taps"] * R * C / tot["loads"], 2 * tot["ffma2"] / ideal,
                tot["cost"] / ideal, 100 * 4 * tot["loads"] / (2 * tot["ffma2"]), tot["chunks"], " ".join(desc)))


def plan_dp(psf, k, R=8, chunk_rows=21, widths=(1, 2), allow4=False):
    """Optimal partition of the sheared columns into groups of the given widths (dynamic programming over columns),
    cost per group = span * step_cost(G) + ceil(span / chunk_rows) * fill_cost(G)."""
    ys, xs = np.nonzero(psf)
    ymin = ys.min()
    xp = xs - k * (ys - ymin)
    x0, x1 = xp.min(), xp.max()
    n = x1 - x0 + 1
    first = np.full(n, 10**6)
    last = np.full(n, -1)
    for y, x in zip(ys, xp - x0):
        first[x] = min(first[x], y)
        last[x] = max(last[x], y)
    kR = R // 2

    def step_cost(G):
        return max(2 * G * kR * C, 4 * (C + G)) + 4

    def fill_cost(G):
        return 4 * (R - 1) * (C + G - 1) + 60

    INF = 10**12
    best = [INF] * (n + 1)
    choice = [0] * (n + 1)
    best[0] = 0
    for i in range(n):
        if best[i] >= INF:
            continue
        for G in widths + ((4,) if allow4 else ()):
            j = min(n, i + G)
            f = first[i:j].min()
            l = last[i:j].max()
            if l < 0:
                c = 0
            else:
                span = l - f + 1
                c = span * step_cost(G) + -(-span // chunk_rows) * fill_cost(G)
            if best[i] + c < best[j]:
                best[j] = best[i] + c
                choice[j] = (i, G, 0 if l < 0 else l - f + 1)
    # walk back
    j = n
    slots = steps = segs = 0
    while j > 0:
        i, G, span = choice[j]
        if span:
            slots += span * G
            steps += span
            segs += 1
        j = i
    return dict(cost=best[n], slots=slots, steps=steps, segs=segs, eff=len(ys) / slots)
