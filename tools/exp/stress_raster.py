"""Randomised stress of the PSF rasteriser against the numpy oracle (bit for bit, fp64): random walks of different step sizes
(a few cells to > 1000 positive cells: the centroid list is flushed in chunks), lengths, exposure fractions (tiny to 1),
canvases 64 / 128 / 256, batch sizes that hit every cluster split.
    python tools/exp/stress_raster.py [seconds] [seed]      (test infrastructure: imports oracle/)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import detectinblur_b200.psf_ops as ops
from oracle import psf_oracle as po


def run(budget=30.0, seed=0):
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n_cases = n_psfs = max_cells = 0
    while time.time() - t0 < budget:
        canvas = int(rng.choice([256, 256, 128, 64]))
        iters = int(rng.choice([2000, 2000, 500, 37, 4096]))
        n = int(rng.choice([1, 2, 5, 19, 40, 70]))
        xs, frs = [], []
        for _ in range(n):
            step = float(rng.choice([0.01, 0.05, 0.2, 0.6, 1.5]))
            d = (rng.standard_normal(iters) + 1j * rng.standard_normal(iters)) * step
            d += (rng.standard_normal() + 1j * rng.standard_normal()) * step * 0.5       # drift
            x = np.cumsum(d) + canvas / 2 * (1 + 1j)
            x = np.clip(x.real, 2.0, canvas - 3.0) + 1j * np.clip(x.imag, 2.0, canvas - 3.0)
            xs.append(x)
            frs.append(float(rng.choice([1.0, 0.5, 1 / 25, 1.5 / iters, 0.999, rng.random()])))
        xs = np.stack(xs)
        frs = np.array(frs)
        raw = ops.rasterize_psfs(xs, frs, "cuda", canvas=canvas, center=False, out_side=canvas, dtype=torch.float64).cpu().numpy()
        cen, offs = ops.rasterize_psfs(xs, frs, "cuda", canvas=canvas, center=True, out_side=canvas, dtype=torch.float64, return_offsets=True)
        cen, offs = cen.cpu().numpy(), offs.cpu().numpy()
        for k in range(n):
            ref = po.rasterize(xs[k], frs[k], canvas)
            max_cells = max(max_cells, int((ref > 0).sum()))
            if not np.array_equal(raw[k], ref):
                print("RAW MISMATCH", dict(case=n_cases, k=k, canvas=canvas, iters=iters, n=n, frac=frs[k], maxdiff=float(np.abs(raw[k] - ref).max())))
                raise AssertionError("rasteriser differs from the oracle")
            ox, oy = po.centroid_offsets(ref, canvas)
            if (int(offs[k][0]), int(offs[k][1])) != (ox, oy) or not np.array_equal(cen[k], po.center(ref, canvas)):
                print("CENTRE MISMATCH", dict(case=n_cases, k=k, canvas=canvas, iters=iters, n=n, frac=frs[k], got=tuple(offs[k]), want=(ox, oy)))
                raise AssertionError("centred PSF differs from the oracle")
        n_cases += 1
        n_psfs += n
    msg = "ok: %d batches, %d PSFs in %.0f s, up to %d positive cells per PSF" % (n_cases, n_psfs, time.time() - t0, max_cells)
    print(msg)
    return msg


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 30.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
