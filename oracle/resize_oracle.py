"""Restatement of the reference's input transform after the blur (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/net_transforms.py: ``normalize`` (:135-139), ``_resize_image_and_masks`` (:36-48: scale
factor from min / max side, then ``torch.nn.functional.interpolate(image[None], scale_factor=s, mode='bilinear',
recompute_scale_factor=True, align_corners=False)``) and ``batch_images`` (:218-249, zero padding to multiples of 32).

The interpolation lives in torch (a third-party dependency of the reference, unpinned; torch 2.11.0 here).  Its published
algorithm for bilinear / align_corners=False / recomputed scale: output size floor(input * scale_factor) evaluated in
double; per axis ``src = fma(in / out, dst + 0.5, -0.5)`` in float32 (the builds contract the expression into one FMA:
verified against torch here, a separately rounded product is off by an ulp of src), clamped below at 0; ``i0 = int(src)``,
``i1 = i0 + (i0 < in - 1)``, ``l1 = src - i0``, ``l0 = 1 - l1``; value = l0h * (l0w * v00 + l1w * v01) + l1h * (l0w * v10 + l1w * v11).
Pinned by tests/golden/resize_cases.npz (outputs of the unmodified reference transform, tools/make_golden.py).
"""
import math

import numpy as np


def resize_scale(h, w, min_size, max_size):
    """net_transforms.py:38-44."""
    lo, hi = float(min(h, w)), float(max(h, w))
    scale = float(min_size) / lo
    if hi * scale > max_size:
        scale = float(max_size) / hi
    return scale


def output_size(h, w, min_size, max_size):
    s = resize_scale(h, w, min_size, max_size)
    return int(math.floor(float(h) * s)), int(math.floor(float(w) * s))


def _axis(n_in, n_out):
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    # fused multiply-add, as torch's CPU and CUDA builds both contract it (the float32 product is exact in float64)
    src = (np.float64(scale) * (dst.astype(np.float64) + 0.5) - 0.5).astype(np.float32)
    src = np.where(src < 0, np.float32(0), src).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), n_in - 1)
    i1 = i0 + (i0 < n_in - 1)
    l1 = np.clip(src - i0.astype(np.float32), 0, 1).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i1, l0, l1


def resize_bilinear(img, out_h, out_w):
    """CHW float32 image -> C x out_h x out_w."""
    img = np.asarray(img, dtype=np.float32)
    y0, y1, ly0, ly1 = _axis(img.shape[1], out_h)
    x0, x1, lx0, lx1 = _axis(img.shape[2], out_w)
    top = img[:, y0][:, :, x0] * lx0 + img[:, y0][:, :, x1] * lx1
    bot = img[:, y1][:, :, x0] * lx0 + img[:, y1][:, :, x1] * lx1
    return (top * ly0[None, :, None] + bot * ly1[None, :, None]).astype(np.float32)


def transform_forward(images, means, stds, min_size, max_size, size_divisible=32):
    """GeneralizedRCNNTransform.forward in eval mode (:82-133) on a list of CHW float32 arrays.

    Returns (zero-padded batch [N, C, Hp, Wp], [(h, w), ...])."""
    outs = []
    for img, mean, std in zip(images, means, stds):
        img = np.asarray(img, dtype=np.float32)
        m = np.asarray(mean, dtype=np.float32)[:, None, None]
        s = np.asarray(std, dtype=np.float32)[:, None, None]
        x = (img - m) / s
        oh, ow = output_size(img.shape[1], img.shape[2], min_size, max_size)
        outs.append(resize_bilinear(x, oh, ow))
    hp = int(math.ceil(float(max(o.shape[1] for o in outs)) / size_divisible) * size_divisible)
    wp = int(math.ceil(float(max(o.shape[2] for o in outs)) / size_divisible) * size_divisible)
    batch = np.zeros((len(outs), outs[0].shape[0], hp, wp), dtype=np.float32)
    for k, o in enumerate(outs):
        batch[k, :, :o.shape[1], :o.shape[2]] = o
    return batch, [(o.shape[1], o.shape[2]) for o in outs]
