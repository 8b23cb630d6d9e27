"""Randomised stress of the tap compaction and the program builders with SYNTHETIC PSFs (random sparse sets, line segments,
dense blobs, taps on the container's border rows / columns, single taps, all-zero), fp32 and fp16 PSFs, sides 128 / 65 / 129 / 256 (the replicate-padded branch):
exact-order kernel against the numpy oracle (bit for bit, fp32), tiled kernels against the exact-order kernel (<= 1e-5).
    python tools/exp/stress_psf.py [seconds] [seed]      (test infrastructure: imports oracle/)"""
import os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import _lib
from oracle import blur_oracle as bo


def run(budget=60.0, seed=0):
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")


    def make_psf(side):
        p = np.zeros((side, side), dtype=np.float32)
        kind = rng.integers(0, 7)
        c = 63
        if kind == 0:       # random sparse set in a random box
            h, w = rng.integers(1, side + 1), rng.integers(1, side + 1)
            y0, x0 = rng.integers(0, side - h + 1), rng.integers(0, side - w + 1)
            n = int(rng.integers(1, 400))
            ys, xs = rng.integers(y0, y0 + h, n), rng.integers(x0, x0 + w, n)
            p[ys, xs] = rng.random(n, dtype=np.float32) + 0.01
        elif kind == 1:     # line segment through the centre region
            L = int(rng.integers(2, 60)); a = rng.random() * np.pi
            t = np.linspace(-L / 2, L / 2, 2 * L)
            ys = np.clip(np.round(c + rng.integers(-8, 9) + t * np.sin(a)).astype(int), 0, side - 1)
            xs = np.clip(np.round(c + rng.integers(-8, 9) + t * np.cos(a)).astype(int), 0, side - 1)
            p[ys, xs] = rng.random(len(t), dtype=np.float32) + 0.01
        elif kind == 2:     # dense blob
            h, w = rng.integers(1, 40), rng.integers(1, 40)
            y0, x0 = rng.integers(max(0, c - 45), min(side - h, c + 10) + 1), rng.integers(max(0, c - 45), min(side - w, c + 10) + 1)
            p[y0:y0 + h, x0:x0 + w] = rng.random((h, w), dtype=np.float32) + 0.01
        elif kind == 3:     # taps on the container's border rows / columns (torch.roll wraps there)
            n = int(rng.integers(1, 20))
            for _ in range(n):
                if rng.random() < 0.5:
                    p[rng.choice([0, side - 1]), rng.integers(0, side)] = rng.random() + 0.01
                else:
                    p[rng.integers(0, side), rng.choice([0, side - 1])] = rng.random() + 0.01
            p[c, c] = 0.5
        elif kind == 4:     # single tap
            p[rng.integers(0, side), rng.integers(0, side)] = 1.0
        elif kind == 5:     # two far-apart clusters (many empty rows / groups between)
            for _ in range(2):
                y0, x0 = rng.integers(1, side - 12), rng.integers(1, side - 12)
                p[y0:y0 + rng.integers(1, 10), x0:x0 + rng.integers(1, 10)] = rng.random() + 0.01
        else:               # ring
            r = rng.integers(3, 50)
            th = np.linspace(0, 2 * np.pi, 6 * r)
            p[np.clip(np.round(c + r * np.sin(th)).astype(int), 0, side - 1), np.clip(np.round(c + r * np.cos(th)).astype(int), 0, side - 1)] = 1.0
        return p


    t0 = time.time()
    n_cases = 0
    stats = {"masked": 0, "dense": 0, "none": 0}
    worst = 0.0
    while time.time() - t0 < budget:
        side = int(rng.choice([128, 128, 128, 65, 129, 256]))
        nb = int(rng.integers(1, 5))
        half_psf = rng.random() < 0.3
        psfs = np.stack([make_psf(side) for _ in range(nb)])
        if half_psf:
            psfs = psfs.astype(np.float16).astype(np.float32)
        if not all(p.sum() > 0 for p in psfs):
            continue
        H, W = int(rng.choice([66, 70, 97, 130])), int(rng.choice([65, 90, 224, 449, 500]))
        imgs_np = [rng.random((int(rng.integers(1, 4)), H, W), dtype=np.float32) for _ in range(nb)]
        imgs = [torch.from_numpy(a).to(dev) for a in imgs_np]
        tpsf = torch.from_numpy(psfs).to(dev)
        ts = ops.compact_taps(tpsf.half() if half_psf else tpsf, normalize=True)
        for m in ts.meta:
            stats["none" if (m.flags & _lib.META_NO_PROGRAM) else ("masked" if m.prog_group_w == 0 else "dense")] += 1
        exact = bf.blur_batch(imgs, ts, list(range(nb)), exact=True)
        tiled = bf.blur_batch(imgs, ts, list(range(nb)))
        for k in range(nb):
            err = float((tiled[k] - exact[k]).abs().max())
            worst = max(worst, err)
            if not err <= 1e-5:
                np.save("/tmp/bad_psf.npy", psfs[k])
                print("TILED MISMATCH", dict(case=n_cases, k=k, side=side, shape=imgs_np[k].shape, err=err, meta=(ts.meta[k].count, ts.meta[k].prog_chunks,
                                                                                                              ts.meta[k].prog_group_w, ts.meta[k].prog_shear)))
                raise AssertionError("mismatch (see the line printed above)")
        # the exact-order kernel against the numpy restatement of the reference loop, one image per batch
        k = int(rng.integers(0, nb))
        pn = bo.normalize_psf(psfs[k].astype(np.float16)).astype(np.float32) if half_psf else bo.normalize_psf(psfs[k])
        want = bo.manual_blur(imgs_np[k], pn)
        got = np.squeeze(exact[k].cpu().numpy())
        want = np.squeeze(np.asarray(want))
        if got.shape != want.shape or not np.array_equal(got, want):
            bad = float(np.abs(got.astype(np.float64) - want).max()) if got.shape == want.shape else -1
            if not (half_psf and bad <= 1e-6):      # a half PSF's sum is rounded to half by torch's own reduction; tolerate the last ulp there
                print("EXACT MISMATCH", dict(case=n_cases, k=k, side=side, shape=imgs_np[k].shape, err=bad, taps=ts.meta[k].count))
                raise AssertionError("mismatch (see the line printed above)")
        n_cases += 1
    msg = ("ok: %d batches in %.0f s, worst tiled-vs-exact %.3g, programs %s" % (n_cases, time.time() - t0, worst, stats))
    print(msg)
    return msg


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
