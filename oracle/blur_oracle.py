"""numpy restatement of the reference's sparse-PSF blur loop (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/models/blur_functions.py:
  * ``normalize_psf``      <- blur_functions.py:98   (``psf_GPU = psf_GPU/psf_GPU.sum()``)
  * ``compact_taps``       <- blur_functions.py:63   (``psf_GPU.nonzero(as_tuple=False)``, row-major)
  * ``source_index``       <- blur_functions.py:17-42 (256 branch) and :43-69 (128 branch): pad, roll, crop
  * ``manual_blur``        <- blur_functions.py:11-89 (loop :66-67, crop :69, noise/clamp :72-74)
  * ``blur_image_list``    <- blur_functions.py:92-100
  * ``normalize_image``    <- models/net_transforms.py:135-139

Arithmetic is restated operation by operation: per tap one rounded multiply and one rounded add in the
image dtype, taps in row-major ``nonzero`` order (no FMA, no reassociation), so fp32 and fp16 results are
bit-identical to the reference loop run on CPU tensors (pinned by tests/golden/blur_*.npz).
"""
import math

import numpy as np

BRANCH_128 = 0  # pad (63, 64), centre 63, reflect (or zeros when H < 64 or W < 64)
BRANCH_256 = 1  # pad (127, 128), centre 127, replicate

PAD_REFLECT = 0
PAD_ZERO = 1
PAD_REPLICATE = 2


def normalize_psf(psf):
    """psf / psf.sum() in the PSF's own dtype (blur_functions.py:98).

    torch accumulates a half sum in fp32 and rounds it to half, then divides half/half through fp32.
    For PSFs whose entries lie on the fp16 grid and sum below 1 (every stored / generated PSF of the
    reference) all partial sums are exact in fp32, so the accumulation order is immaterial.
    """
    psf = np.asarray(psf)
    if psf.dtype == np.float16:
        s = np.float16(np.float32(psf.astype(np.float64).sum()))
        return (psf.astype(np.float32) / np.float32(s)).astype(np.float16)
    if psf.dtype == np.float32:
        s = np.float32(psf.astype(np.float64).sum())
        return (psf / s).astype(np.float32)
    s = psf.sum()
    return psf / s


def compact_taps(psf_normalized):
    """Row-major nonzero taps of the (already normalised) PSF: (ys, xs, weights).

    blur_functions.py:63 -- ``nonzero`` enumerates in row-major order, which is the accumulation order.
    """
    psf = np.asarray(psf_normalized)
    ys, xs = np.nonzero(psf)
    return ys.astype(np.int32), xs.astype(np.int32), psf[ys, xs]


def branch_of(psf_side):
    """blur_functions.py:17 -- ``psf.shape[0] > 129`` selects the 256 branch."""
    return BRANCH_256 if psf_side > 129 else BRANCH_128


def pad_mode_of(branch, H, W):
    """blur_functions.py:28-31 (always replicate) and :55-58 (zeros below 64, else reflect)."""
    if branch == BRANCH_256:
        return PAD_REPLICATE
    if H < 64 or W < 64:
        return PAD_ZERO
    if H <= 64 or W <= 64:
        # torch reflect padding needs pad < dim; the reference raises here (pad 64 on a 64-px side).
        raise RuntimeError("reflect padding of 64 needs every image side > 64 (reference raises too)")
    return PAD_REFLECT


def source_index(n, coord, branch, pad_mode):
    """Source row (or column) read by output positions 0..n-1 for a tap at PSF row (column) ``coord``.

    Restates pad -> torch.roll(shift = coord - centre) -> crop:
      padded length  L = n + lo + hi           (lo, hi) = (63, 64) | (127, 128)
      out[i] = padded[(i + lo - (coord - lo)) mod L]
      padded[q] = img[map(q - lo)]
    -1 marks a zero-padded read.
    """
    lo, hi = (127, 128) if branch == BRANCH_256 else (63, 64)
    L = n + lo + hi
    i = np.arange(n, dtype=np.int64)
    q = (i + 2 * lo - int(coord)) % L
    s = q - lo
    if pad_mode == PAD_REFLECT:
        s = np.where(s < 0, -s, s)
        s = np.where(s >= n, 2 * (n - 1) - s, s)
    elif pad_mode == PAD_REPLICATE:
        s = np.clip(s, 0, n - 1)
    else:
        s = np.where((s < 0) | (s >= n), -1, s)
    return s


def manual_blur(image, psf, noise=None, noise_var=None, pad_mode=None):
    """blur_functions.py:11-74 on a CHW numpy image (float32 or float16); psf is already normalised.

    ``noise`` (same shape as the result) and ``noise_var`` restate :72-74 with the random draws supplied
    by the caller: ``clamp(out + noise * sqrt(noise_var), 0, 1)``.  ``pad_mode`` overrides the reference's own choice
    of boundary (:55-58); the Fourier-path mirror uses zero padding at sizes where manual_blur itself would reflect.
    Returns a CxHxW array (HxW when C == 1, mirroring the ``.squeeze()`` at :69).
    """
    image = np.asarray(image)
    assert image.ndim == 3
    C, H, W = image.shape
    psf = np.asarray(psf)
    branch = branch_of(psf.shape[0])
    pad_mode = pad_mode_of(branch, H, W) if pad_mode is None else pad_mode
    ys, xs, ws = compact_taps(psf)
    dt = image.dtype.type
    out = np.zeros((C, H, W), dtype=image.dtype)
    zero = dt(0)
    for y, x, w in zip(ys, xs, ws):
        rows = source_index(H, y, branch, pad_mode)
        cols = source_index(W, x, branch, pad_mode)
        g = image[:, np.maximum(rows, 0)[:, None], np.maximum(cols, 0)[None, :]]
        if pad_mode == PAD_ZERO:
            mask = (rows[:, None] < 0) | (cols[None, :] < 0)
            g = np.where(mask[None], zero, g)
        # the 0-dim PSF element is cast to the image dtype by torch's type promotion
        out = (out + (g * dt(w)).astype(image.dtype)).astype(image.dtype)
    if noise is not None:
        sd = math.sqrt(noise_var)
        out = np.clip((out + (np.asarray(noise, dtype=image.dtype) * dt(sd)).astype(image.dtype)).astype(image.dtype),
                      dt(0), dt(1))
    if C == 1:
        out = out[0]
    return out


def blur_image_list(images, blur_dicts, psfs):
    """blur_functions.py:92-100: in-place over the list, skipping entries whose blur_dict['blurring'] is falsy."""
    for idx, (img, bd, psf) in enumerate(zip(images, blur_dicts, psfs)):
        if not bd["blurring"]:
            continue
        images[idx] = manual_blur(img, normalize_psf(psf))


def normalize_image(image, mean, std):
    """models/net_transforms.py:135-139: (image - mean[:,None,None]) / std[:,None,None] in the image dtype."""
    image = np.asarray(image)
    m = np.asarray(mean).astype(image.dtype)
    s = np.asarray(std).astype(image.dtype)
    return ((image - m[:, None, None]).astype(image.dtype) / s[:, None, None]).astype(image.dtype)


def tap_extents(psf_normalized):
    """utils.py:372-380 (expand_targets): min/max tap offsets relative to centre 63 -> (left, top, right, bottom)."""
    ys, xs, _ = compact_taps(psf_normalized)
    return int(xs.min()) - 63, int(ys.min()) - 63, int(xs.max()) - 63, int(ys.max()) - 63


def psf_pca(psf):
    """transforms.py:366-385: second moments of the PSF support -> (theta_rad, lambda1 scale, lambda2 scale)."""
    ys, xs = np.nonzero(np.asarray(psf) > 0)
    yp = ys - ys.mean()
    xp = xs - xs.mean()
    cov = (yp * xp).mean()
    var_x = (xp * xp).mean()
    var_y = (yp * yp).mean()
    root = math.sqrt(math.pow((var_x - var_y) / 2, 2) + math.pow(cov, 2))
    lam1 = (var_x + var_y) / 2 + root
    lam2 = (var_x + var_y) / 2 - root

    def sigmoid(v):
        return 1 / (1 + math.exp(-v))

    s1 = 1 - (sigmoid(math.sqrt(lam1) / 10) - 0.5) * 0.6
    s2 = 1 - (sigmoid(math.sqrt(lam2) / 10) - 0.5) * 0.6
    theta = -math.atan2(lam1 - var_x, -cov)
    return theta, s1, s2
