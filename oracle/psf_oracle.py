"""numpy restatement of the reference's trajectory -> PSF path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows:
  * ``trajectory``     <- /root/reference/motion_blur/generate_trajectory.py:38-98 (Boracchi-Foi random walk,
                          numpy global MT19937 stream consumed in the same order)
  * ``trajectory_philox`` the same walk driven by a counter-based generator: what the GPU generator computes
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


def philox4x32(seed, ctr_lo, ctr_hi):
    """Philox4x32-10 (Salmon et al., SC'11; the generator family behind torch's CUDA randn): key = the 64-bit seed,
    counter = (ctr_lo, ctr_hi) as two 64-bit halves.  Returns the four 32-bit output words."""
    m32 = 0xFFFFFFFF
    c = [ctr_lo & m32, (ctr_lo >> 32) & m32, ctr_hi & m32, (ctr_hi >> 32) & m32]
    k0, k1 = seed & m32, (seed >> 32) & m32
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & m32, p1 & m32, ((p0 >> 32) ^ c[3] ^ k1) & m32, p0 & m32]
        k0, k1 = (k0 + 0x9E3779B9) & m32, (k1 + 0xBB67AE85) & m32
    return c


def trajectory_philox(seed, index, expl, canvas=256, iters=2000, max_len=96):
    """The counter-based variant of ``trajectory`` that dib_generate_trajectories evaluates (csrc/trajectory.cu): the
    update equations of generate_trajectory.py:38-98 with the randomness of step t taken from
    Philox(seed; counter t + 1, stream index): word 0 = shake test, word 1 = shake angle, words 2-3 = Box-Muller pair;
    counter 0 = the four shape parameters.  Returns (x complex128[iters], number of impulsive shakes)."""
    def u01(r):
        return (float(r) + 0.5) * (1.0 / 4294967296.0)

    r0 = philox4x32(seed, 0, index)
    centripetal, prob_big_shake = 0.7 * u01(r0[0]), 0.2 * u01(r0[1])
    gaussian_shake, ang = 10.0 * u01(r0[2]), 360.0 * u01(r0[3]) * (np.pi / 180.0)
    step = max_len / float(iters - 1)
    scale = expl if expl > 0 else step
    vx, vy = np.cos(ang) * scale, np.sin(ang) * scale
    x = y = 0.0
    out = np.zeros(iters, dtype=np.complex128)
    big = 0
    for t in range(iters - 1):
        r = philox4x32(seed, t + 1, index)
        kx = ky = 0.0
        if u01(r[0]) < prob_big_shake * expl:
            a = np.pi + (u01(r[1]) - 0.5)
            c, s = np.cos(a), np.sin(a)
            kx, ky = 2.0 * (vx * c - vy * s), 2.0 * (vx * s + vy * c)
            big += 1
        rad = np.sqrt(-2.0 * np.log(u01(r[2])))
        a = 6.28318530717958647692 * u01(r[3])
        g0, g1 = rad * np.cos(a), rad * np.sin(a)
        vx += kx + expl * (gaussian_shake * g0 - centripetal * x) * step
        vy += ky + expl * (gaussian_shake * g1 - centripetal * y) * step
        inv = step / np.sqrt(vx * vx + vy * vy)
        vx, vy = vx * inv, vy * inv
        x, y = x + vx, y + vy
        out[t + 1] = complex(x, y)
    return out + complex(canvas / 2, canvas / 2), big


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
