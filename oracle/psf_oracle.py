"""numpy restatement of the reference's trajectory -> PSF path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows:
  * ``trajectory``     <- /root/reference/motion_blur/generate_trajectory.py:38-98 (Boracchi-Foi random walk,
                          numpy global MT19937 stream consumed in the same order)
  * ``time_weights``   <- motion_blur/generate_PSF.py:47-56 (exposure-fraction time slicing, one fraction)
  * ``rasterize``      <- motion_blur/generate_PSF.py:31-77 (4-corner bilinear splat, sequential fp64 sums, /iters)
  * ``center``         <- motion_blur/generate_PSF.py:106-123 (weighted centroid, int() truncation, np.roll)
  * ``stored_psf``     <- dataset_utils/generate_PSFs.py:47-60 (Trajectory.fit().fit(), PSF.fit(), centerPSF, fp16)
  * ``crop128``        <- transforms.py:308-309 / :334-335

All arithmetic is float64 and order-preserving: each PSF cell receives its contributions in ascending
sample order exactly as the reference's Python loop does.
"""
import numpy as np


def trajectory(canvas=256, iters=2000, max_len=96, expl=0.005, rng=np.random):
    """generate_trajectory.py:38-98.  Returns x (complex128[iters]) already shifted to the canvas centre.

    ``rng`` must expose uniform()/randn() like numpy's global module (the reference uses the global stream).
    """
    centripetal = 0.7 * rng.uniform(0, 1)
    prob_big_shake = 0.2 * rng.uniform(0, 1)
    gaussian_shake = 10 * rng.uniform(0, 1)
    init_angle = 360 * rng.uniform(0, 1)
    v0 = complex(real=np.cos(np.deg2rad(init_angle)), imag=np.sin(np.deg2rad(init_angle)))
    step = max_len / (iters - 1)
    v = v0 * max_len / (iters - 1)
    if expl > 0:
        v = v0 * expl
    x = np.array([complex(real=0, imag=0)] * iters)
    for t in range(0, iters - 1):
        if rng.uniform() < prob_big_shake * expl:
            next_direction = 2 * v * (np.exp(complex(real=0, imag=np.pi + (rng.uniform() - 0.5))))
        else:
            next_direction = 0
        dv = next_direction + expl * (
            gaussian_shake * complex(real=rng.randn(), imag=rng.randn()) - centripetal * x[t]) * step
        v += dv
        v = (v / float(np.abs(v))) * (max_len / float((iters - 1)))
        x[t + 1] = x[t] + v
    return x + complex(canvas / 2, canvas / 2)


def time_weights(iters, fraction):
    """generate_PSF.py:47-56 for a single fraction (prevT = 0): weight of every trajectory sample."""
    w = np.zeros(iters, dtype=np.float64)
    fi = fraction * iters
    prev = 0 * iters
    for t in range(iters):
        if (fi >= t) and (prev < t - 1):
            w[t] = 1
        elif (fi >= t - 1) and (prev < t - 1):
            w[t] = fi - (t - 1)
        elif (fi >= t) and (prev < t):
            w[t] = t - prev
        elif (fi >= t - 1) and (prev < t):
            w[t] = (fraction - 0) * iters
        else:
            w[t] = 0
    return w


def _tri(v):
    return np.maximum(0, (1 - np.abs(v)))


def rasterize(x, fraction, canvas=256):
    """generate_PSF.py:31-77: PSF canvas (float64) of one exposure fraction from trajectory samples x."""
    x = np.asarray(x, dtype=np.complex128)
    iters = len(x)
    w = time_weights(iters, fraction)
    psf = np.zeros((canvas, canvas), dtype=np.float64)
    re, im = x.real, x.imag
    m2 = np.minimum(canvas - 1, np.maximum(1, np.floor(re))).astype(np.int64)
    m1 = np.minimum(canvas - 1, np.maximum(1, np.floor(im))).astype(np.int64)
    M2, M1 = m2 + 1, m1 + 1
    # corner order per sample as in generate_PSF.py:64-75: (m1,m2) (m1,M2) (M1,m2) (M1,M2)
    rows = np.stack([m1, m1, M1, M1], axis=1).ravel()
    cols = np.stack([m2, M2, m2, M2], axis=1).ravel()
    vals = np.stack([
        w * (_tri(re - m2) * _tri(im - m1)),
        w * (_tri(re - M2) * _tri(im - m1)),
        w * (_tri(re - m2) * _tri(im - M1)),
        w * (_tri(re - M2) * _tri(im - M1)),
    ], axis=1).ravel()
    # np.add.at applies the updates one by one in index order == ascending t then corner, which is the
    # reference loop's accumulation order for every cell (fp64 sums are order-sensitive).
    # canvas-1 clamp + 1 can address row/col == canvas only if the trajectory leaves the canvas, where the
    # reference raises IndexError as well.
    np.add.at(psf, (rows, cols), vals)
    return psf / iters


def centroid_offsets(psf, canvas=256):
    """generate_PSF.py:106-120: (offsetX, offsetY) = int(weighted centroid - canvas/2), row-major accumulation."""
    total = np.sum(psf)
    ys, xs = np.nonzero(psf > 0)
    ax = 0.0
    ay = 0.0
    for cx, cy in zip(xs, ys):
        weight = psf[cy, cx] / total
        ax += cx * weight
        ay += cy * weight
    return int(ax - canvas / 2), int(ay - canvas / 2)


def center(psf, canvas=256):
    """generate_PSF.py:106-123."""
    ox, oy = centroid_offsets(psf, canvas)
    psf = np.roll(psf, shift=-ox, axis=1)
    psf = np.roll(psf, shift=-oy, axis=0)
    return psf


def crop128(psf):
    """transforms.py:308-309: the 256 canvas keeps its central 128 window."""
    if psf.shape[0] > 128:
        return psf[64:128 + 64, 64:128 + 64]
    return psf


def stored_psf(expl, fraction, rng=np.random, canvas=256, max_len=96):
    """dataset_utils/generate_PSFs.py:47-60: what one file of the stored bank holds (float16 256x256).

    The reference calls ``Trajectory(...).fit().fit()``: two trajectories are drawn, the second is used.
    """
    trajectory(canvas, 2000, max_len, expl, rng)
    x = trajectory(canvas, 2000, max_len, expl, rng)
    return center(rasterize(x, fraction, canvas), canvas).astype(np.float16), x
