"""Restatement of the reference's CPU Fourier-domain blur (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

Follows /root/reference/motion_blur/blur_image.py:25-154 (``BlurImageHandler.__init__`` + ``blur_image``):
edge-pad by half the kernel (:77-85), zero-pad the PSF to the image (:119-123), min-max normalise PSF and
image (:128-130), one ``scipy.signal.fftconvolve(..., 'same')`` per channel (:131-133), min-max normalise the
result (:134), unpad (:137-140), uint8 truncation (:147).

Third-party arithmetic the reference relies on (not vendored under /root/reference, versions unpinned there;
container versions): ``scipy.signal.fftconvolve`` (scipy 1.18.1, used directly here as the reference does) and
``cv2.normalize(NORM_MINMAX, CV_32F)`` (OpenCV 4.13.0), restated in numpy as ``src * scale + shift`` with
``scale = 1 / (max - min)``, ``shift = -min * scale`` evaluated in float32 like OpenCV's convertTo.

This is the ``--cpu_blur`` path; it is timed as the reported CPU baseline (bench.py) and is NOT what the CUDA
path is bit-compared with: it differs from the GPU loop by boundary handling (edge vs reflect), the final
per-image min-max contrast stretch and uint8 truncation.
"""
import math

import numpy as np
from scipy import signal


def minmax_normalize_f32(a):
    """cv2.normalize(a, a, alpha=0, beta=1, norm_type=NORM_MINMAX, dtype=CV_32F) (blur_image.py:128-134)."""
    a = np.asarray(a)
    smin = float(a.min())
    smax = float(a.max())
    d = smax - smin
    scale = (1.0 / d) if d > np.finfo(np.float64).eps else 0.0
    shift = 0.0 - smin * scale
    if a.dtype == np.float32:
        return a * np.float32(scale) + np.float32(shift)
    return (a.astype(np.float64) * scale + shift).astype(np.float32)


def _prepare(image, psf):
    """BlurImageHandler.__init__ (blur_image.py:55-97): optional bicubic upscale, edge padding, grey -> RGB.

    Returns (padded HxWx3 array, (pr, pc) padding, original PIL size or None)."""
    from PIL import Image
    pil = image if isinstance(image, Image.Image) else Image.fromarray(np.asarray(image))
    key, kex = psf.shape
    original_size = pil.size
    yN, xN = pil.size          # PIL .size is (W, H): the reference compares W with the kernel rows and H with its cols (:57-61)
    if yN - key < 0 or xN - kex < 0:
        ratio_y, ratio_x = key / yN, kex / xN
        r = ratio_x if ratio_x > ratio_y else ratio_y
        pil = pil.resize((math.ceil(r * xN), math.ceil(r * yN)), Image.BICUBIC)     # (:65/:67: the swapped names are the reference's)
    else:
        original_size = None
    img = np.array(pil)
    pr, pc = round(key / 2), round(kex / 2)
    pad = ((pr, pr), (pc, pc), (0, 0)) if img.ndim > 2 else ((pr, pr), (pc, pc))
    img = np.pad(img, pad_width=pad, mode="edge")
    if img.ndim < 3:
        img = np.stack([img] * 3, axis=2).astype(np.float64)      # :91-97 builds the RGB copy with np.zeros (float64)
    return img, (pr, pc), original_size


def _finish(blurred, pads, original_size):
    """blur_image.py:134-147: min-max stretch, unpad, optional Lanczos resize back, abs / uint8 truncation."""
    pr, pc = pads
    blurred = minmax_normalize_f32(blurred)
    blurred = blurred[pr:blurred.shape[0] - pr, pc:blurred.shape[1] - pc, :]
    if original_size is not None:
        import cv2
        blurred = cv2.resize(blurred, original_size, interpolation=cv2.INTER_LANCZOS4)
    return np.abs(blurred), (blurred * 255).astype(np.uint8)


def padded_kernel(psf, yN, xN, old_delta_pad=False):
    """The kernel zero-padded to the padded image (blur_image.py:114-123)."""
    key, kex = psf.shape
    dY, dX = yN - key, xN - kex
    if old_delta_pad:
        return np.pad(psf, dX // 2, "constant")
    return np.pad(psf, ((dY // 2, math.ceil(dY / 2)), (math.ceil(dX / 2), dX // 2)), "constant")


def fourier_blur(image, psf, old_delta_pad=False):
    """blur_image.py:25-154 for a PIL image or an HxWx3 / HxW uint8 array.

    Returns (float32 HxWx3 result == ``handler.result[0]``, uint8 HxWx3 == ``np.array(handler.pilImageResult)``).
    """
    psf = np.asarray(psf, dtype=np.float32)
    padded, pads, original_size = _prepare(image, psf)
    yN, xN, _ = padded.shape
    tmp = minmax_normalize_f32(padded_kernel(psf, yN, xN, old_delta_pad))
    blurred = minmax_normalize_f32(padded)
    for ch in range(3):
        blurred[:, :, ch] = np.array(signal.fftconvolve(blurred[:, :, ch], tmp, "same"))
    return _finish(blurred, pads, original_size)


def kernel_centre(psf_shape, yN, xN, old_delta_pad=False):
    """Where ``fftconvolve(image, padded_kernel, 'same')`` puts the kernel's origin: output (i, j) reads the image at
    (i + cy - y, j + cx - x) for kernel element (y, x), zero outside the padded image.  'same' keeps the centre
    (n - 1) // 2 of the full convolution and the kernel sits at (top, left) of its zero-padded canvas, so
    cy = (yN - 1) // 2 - top, cx = (xN - 1) // 2 - left: (63, 63) for even padded sizes, 64 rows for an odd height."""
    key, kex = psf_shape
    dY, dX = yN - key, xN - kex
    if old_delta_pad:
        top = left = dX // 2
        ty, tx = key + 2 * top, kex + 2 * left
        # the kernel canvas need not match the image here; 'same' is centred w.r.t. the full output of image (*) canvas
        return (ty - 1) // 2 - top, (tx - 1) // 2 - left
    return (yN - 1) // 2 - dY // 2, (xN - 1) // 2 - math.ceil(dX / 2)


def spatial_blur(image, psf, old_delta_pad=False):
    """The same result as ``fourier_blur`` written as the tap sum the CUDA path evaluates: zero-boundary convolution
    of the edge-padded, min-max normalised image with ``psf / max`` about ``kernel_centre``, then the shared finish."""
    psf = np.asarray(psf, dtype=np.float32)
    padded, pads, original_size = _prepare(image, psf)
    yN, xN, _ = padded.shape
    cy, cx = kernel_centre(psf.shape, yN, xN, old_delta_pad)
    w = minmax_normalize_f32(padded_kernel(psf, yN, xN, old_delta_pad))
    if float(w.min()) != 0.0:
        raise ValueError("kernel without zero entries: the min-max normalised canvas is dense")
    kmax = float(psf.max())
    src = minmax_normalize_f32(padded).astype(np.float64)
    acc = np.zeros_like(src)
    ys, xs = np.nonzero(psf)
    for y, x in zip(ys, xs):
        wt = float(np.float32(psf[y, x]) * np.float32(1.0 / kmax))
        oy, ox = cy - int(y), cx - int(x)           # acc[i, j] += wt * src[i + oy, j + ox]
        i0, i1 = max(0, -oy), min(yN, yN - oy)
        j0, j1 = max(0, -ox), min(xN, xN - ox)
        if i0 < i1 and j0 < j1:
            acc[i0:i1, j0:j1] += wt * src[i0 + oy:i1 + oy, j0 + ox:j1 + ox]
    return _finish(acc.astype(np.float32), pads, original_size)
