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


def fourier_blur(image_hwc_u8, psf):
    """blur_image.py:25-154 for an HxWx3 uint8 image no smaller than the kernel.

    Returns (float32 HxWx3 result == ``handler.result[0]``, uint8 HxWx3 == ``np.array(handler.pilImageResult)``).
    """
    img = np.asarray(image_hwc_u8)
    psf = np.asarray(psf, dtype=np.float32)
    key, kex = psf.shape
    if img.ndim == 2:
        img = np.stack([img] * 3, axis=2)
    # PIL .size is (W, H); the reference compares W with the kernel rows and H with the kernel cols (:57-61)
    if img.shape[1] - key < 0 or img.shape[0] - kex < 0:
        raise NotImplementedError("images smaller than the kernel take the bicubic-upscale branch (:62-69); "
                                  "not needed for the baseline")
    pr, pc = round(key / 2), round(kex / 2)
    padded = np.pad(img, ((pr, pr), (pc, pc), (0, 0)), mode="edge")
    yN, xN, _ = padded.shape
    dY, dX = yN - key, xN - kex
    tmp = np.pad(psf, ((dY // 2, math.ceil(dY / 2)), (math.ceil(dX / 2), dX // 2)), "constant")
    tmp = minmax_normalize_f32(tmp)
    blurred = minmax_normalize_f32(padded)
    for ch in range(3):
        blurred[:, :, ch] = np.array(signal.fftconvolve(blurred[:, :, ch], tmp, "same"))
    blurred = minmax_normalize_f32(blurred)
    blurred = blurred[pr:blurred.shape[0] - pr, pc:blurred.shape[1] - pc, :]
    return np.abs(blurred), (blurred * 255).astype(np.uint8)
