"""Mirror of the input-transform step of the reference's ``models/net_transforms.py`` that follows the blur.

``GeneralizedRCNNTransform.forward`` (net_transforms.py:82-133) normalises every image with its own row of
``newMeans`` / ``newSTDs`` (``normalize`` :135-139), resizes (:151-175) and packs the list into one zero-padded batch
whose sides are multiples of 32 (``batch_images`` :218-249).  Here

  * for images already at the requested scale (800x1333) the normalisation is the epilogue of the blur kernel and the
    result lands directly inside the padded batch tensor (``fused_blur_normalize``): blurred pixels are written once;
  * for images that need resizing (native-size COCO images) one more kernel, ``dib_resize_batch``, does normalize +
    bilinear resize + zero padding in a single pass over the batch (``resize_normalize_batch``), instead of the
    reference's five passes; ``fused_blur_normalize(..., min_size=, max_size=)`` chains it after the blur.

A model built with ``GeneralizedRCNNTransform(..., normalize_images=False)`` -- the hook the reference already has
(:70-80, :112-118) -- consumes the batch unchanged (``forward`` passes an ``ImageList`` through).
"""
import ctypes
import math

import torch

from . import _lib
from . import blur_functions
from . import psf_ops

CANONICAL_MEAN = [0.485, 0.456, 0.406]      # utils.py:221
CANONICAL_STD = [0.229, 0.224, 0.225]       # utils.py:222


class ImageList(object):
    """The two fields of torchvision's ImageList that the detector reads (net_transforms.py:129-131)."""

    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, device):
        return ImageList(self.tensors.to(device), self.image_sizes)


def normalize(image, image_mean, image_std):
    """net_transforms.py:135-139."""
    dtype, device = image.dtype, image.device
    mean = torch.as_tensor(image_mean, dtype=dtype, device=device)
    std = torch.as_tensor(image_std, dtype=dtype, device=device)
    return (image - mean[:, None, None]) / std[:, None, None]


def resize_scale(h, w, min_size, max_size):
    """Scale factor of GeneralizedRCNNTransform.resize for one image (net_transforms.py:36-48, eval / fixed min_size)."""
    lo, hi = float(min(h, w)), float(max(h, w))
    scale = float(min_size) / lo
    if hi * scale > max_size:
        scale = float(max_size) / hi
    return scale


def padded_batch_shape(sizes, size_divisible=32):
    """batch_images' padded extent (net_transforms.py:236-247)."""
    hmax = max(s[0] for s in sizes)
    wmax = max(s[1] for s in sizes)
    hp = int(math.ceil(float(hmax) / size_divisible) * size_divisible)
    wp = int(math.ceil(float(wmax) / size_divisible) * size_divisible)
    return hp, wp


_launch_count = 0


def launch_count():
    """Kernels launched by this module (resize passes); the blur's own count is blur_functions.launch_count()."""
    return _launch_count


def resize_normalize_batch(images, min_size, max_size, means=None, stds=None, size_divisible=32):
    """normalize + resize + batch_images (net_transforms.py:82-133, eval-mode resize) of CHW CUDA tensors in ONE pass.

    images      list of [C, H, W] CUDA tensors (float32 or float16, any sizes, rows contiguous)
    min_size    the short-side target: one number, or one per image (training draws it per image, :151-158)
    means/stds  per-image (C,) sequences, or None to skip the normalisation (``normalize_images=False``)
    Returns ImageList(tensors=[N, C, Hp, Wp], image_sizes=[(h, w), ...]) with every element of the batch written here.
    """
    global _launch_count
    n = len(images)
    if n == 0:
        raise ValueError("empty image list")
    dev, dtype = images[0].device, images[0].dtype
    if dtype not in (torch.float32, torch.float16):
        raise TypeError("resize_normalize_batch takes float32 or float16 images, got %s" % dtype)
    srcs, sizes = [], []
    for k, img in enumerate(images):
        psf_ops._require_cuda(img, "image %d" % k)
        if img.dim() != 3:
            raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W], got {}".format(img.shape))
        if img.dtype != dtype or img.device != dev or img.shape[0] != images[0].shape[0]:
            raise TypeError("all images of a batch must share dtype, device and channel count")
        if img.stride(2) != 1:
            img = img.contiguous()
        srcs.append(img)
        h, w = int(img.shape[1]), int(img.shape[2])
        scale = resize_scale(h, w, min_size[k] if isinstance(min_size, (list, tuple)) else min_size, max_size)
        sizes.append((int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))))   # interpolate's output size
    hp, wp = padded_batch_shape(sizes, size_divisible)
    C = int(images[0].shape[0])
    batch = torch.empty((n, C, hp, wp), dtype=dtype, device=dev)
    launches = ctypes.c_int(0)
    with psf_ops._on_device(dev):
        stream = psf_ops._stream_ptr(dev)
        for lo in range(0, n, _lib.MAX_BATCH):
            cnt = min(_lib.MAX_BATCH, n - lo)
            descs = (_lib.ResizeImage * cnt)()
            for j in range(cnt):
                k = lo + j
                d, img = descs[j], srcs[k]
                d.src, d.dst = img.data_ptr(), batch[k].data_ptr()
                d.C, d.in_h, d.in_w = C, int(img.shape[1]), int(img.shape[2])
                d.out_h, d.out_w = sizes[k]
                d.pad_h, d.pad_w = hp, wp
                d.src_row_pitch, d.src_chan_pitch = img.stride(1), img.stride(0)
                d.dst_row_pitch, d.dst_chan_pitch = batch.stride(2), batch.stride(1)
                d.normalize = 0
                if means is not None and means[k] is not None:
                    d.normalize = 1
                    for c in range(min(C, 4)):
                        d.mean[c] = float(means[k][c])
                        d.std[c] = float(stds[k][c])
            _lib.check(_lib.lib.dib_resize_batch(descs, cnt, blur_functions._DT[dtype], ctypes.byref(launches), stream))
            _launch_count += launches.value
    return ImageList(batch, sizes)


def fused_blur_normalize(images_GPU, blur_dicts, psfs_GPU, newMeans=None, newSTDs=None, image_mean=None, image_std=None,
                         size_divisible=32, exact=None, min_size=None, max_size=None):
    """blur_image_list + normalize + batch_images in one pass over the pixels.

    images_GPU   list of [3, H, W] CUDA tensors (float32), as handed to blur_image_list
    blur_dicts   per image; entries with a falsy "blurring" are only normalised
    psfs_GPU     per image dense PSF (ignored where not blurring)
    newMeans / newSTDs   optional [N, 3] per-image statistics (utils.get_norm_params, utils.py:219-273); the
                 canonical ImageNet statistics otherwise
    min_size / max_size  the transform's resize target (eval mode: the last of ``min_size``).  When every image is
                 already at that scale -- or no target is given -- the blur kernel normalises and writes straight into
                 the batch; otherwise the images are blurred at native size and ``resize_normalize_batch`` finishes.
    Returns ImageList(tensors=[N, 3, Hp, Wp] zero padded, image_sizes=[(H, W), ...]).
    """
    n = len(images_GPU)
    if n == 0:
        raise ValueError("empty image list")
    image_mean = CANONICAL_MEAN if image_mean is None else image_mean
    image_std = CANONICAL_STD if image_std is None else image_std
    dev, dtype = images_GPU[0].device, images_GPU[0].dtype
    sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images_GPU]
    means = [list(newMeans[i]) if newMeans is not None else list(image_mean) for i in range(n)]
    stds = [list(newSTDs[i]) if newSTDs is not None else list(image_std) for i in range(n)]
    blurred = [k for k in range(n) if blur_dicts[k]["blurring"]]
    idx = [-1] * n
    tapset = None
    if blurred:
        sides = {int(psfs_GPU[k].shape[0]) for k in blurred}
        if len(sides) != 1:
            raise ValueError("fused_blur_normalize needs PSFs of one container size per batch")
        for k in blurred:
            blur_functions.pad_mode_for(int(psfs_GPU[k].shape[0]), sizes[k][0], sizes[k][1])
        psf_dtypes = {psfs_GPU[k].dtype for k in blurred}
        if psf_dtypes == {dtype}:
            tapset = psf_ops.compact_taps(torch.stack([psfs_GPU[k] for k in blurred]), normalize=True)
        else:
            # the reference normalises in the PSF's dtype and multiplies by the 0-dim element cast to the image dtype
            # (blur_functions.py:98, :67): normalise first, cast, then compact -- as blur_image_list does
            dense = torch.stack([(psfs_GPU[k] / psfs_GPU[k].sum()).to(dtype) for k in blurred])
            tapset = psf_ops.compact_taps(dense, normalize=False)
        for j, k in enumerate(blurred):
            idx[k] = j
    if min_size is not None:
        target = min_size[-1] if isinstance(min_size, (list, tuple)) else min_size
        if any(resize_scale(h, w, target, max_size) != 1.0 for h, w in sizes):
            blurred = blur_functions.blur_batch(list(images_GPU), tapset, idx, exact=exact) if tapset is not None else list(images_GPU)
            return resize_normalize_batch(blurred, target, max_size, means, stds, size_divisible)
    hp, wp = padded_batch_shape(sizes, size_divisible)
    batch = torch.zeros((n, int(images_GPU[0].shape[0]), hp, wp), dtype=dtype, device=dev)
    outs = [batch[k, :, :sizes[k][0], :sizes[k][1]] for k in range(n)]
    blur_functions.blur_batch(list(images_GPU), tapset, idx, outs=outs, mean=means, std=stds, exact=exact)
    return ImageList(batch, sizes)


class GeneralizedRCNNTransform(torch.nn.Module):
    """net_transforms.py:58-133 for the blur path: same constructor and ``forward`` contract.

    ``forward`` normalises (unless ``normalize_images`` is False, e.g. after ``fused_blur_normalize``), resizes with
    bilinear interpolation when the scale is not 1, and zero-pads into a batch.  Target resizing and the
    ``crop_images`` training augmentation belong to the detector and are left to the reference code.
    """

    def __init__(self, min_size, max_size, image_mean, image_std, crop_images=False, training=True, normalize_images=True):
        super(GeneralizedRCNNTransform, self).__init__()
        if not isinstance(min_size, (list, tuple)):
            min_size = (min_size,)
        self.min_size = min_size
        self.max_size = max_size
        self.image_mean = image_mean
        self.image_std = image_std
        if crop_images:
            # batch_images' crop branch (net_transforms.py:218-236: crop every image to the batch's smallest extent) is a
            # training augmentation of the detector; silently padding instead would change the batch shapes
            raise NotImplementedError("crop_images=True is not mirrored here; use the reference's transform for that augmentation")
        self.crop_images = crop_images
        self.training = training
        self.normalize_images = normalize_images

    def normalize(self, image, image_mean, image_std):
        return normalize(image, image_mean, image_std)

    def forward(self, images, targets=None, newMeans=None, newSTDs=None):
        if isinstance(images, ImageList):
            return images, targets                   # already normalised and batched by fused_blur_normalize
        images = [img for img in images]
        if images and all(img.is_cuda and img.dim() == 3 and img.dtype in (torch.float32, torch.float16) for img in images):
            # CUDA tensors: normalize + resize + batch in one kernel pass.  Training draws the short-side target per image
            # from torch's CPU generator exactly as torch_choice does (:141-149, :154-155), so seeded runs stay in step;
            # target (box / mask / keypoint) resizing stays with the reference's own transform.
            if targets is not None:
                raise NotImplementedError("target resizing is the detector's side of the transform; pass targets=None here")
            if self.training:
                target_sizes = [float(self.min_size[int(torch.empty(1).uniform_(0., float(len(self.min_size))).item())])
                                for _ in images]
            else:
                target_sizes = [float(self.min_size[-1])] * len(images)
            means = stds = None
            if self.normalize_images:
                if newMeans is not None:
                    means, stds = [newMeans[i, :] for i in range(len(images))], [newSTDs[i, :] for i in range(len(images))]
                else:
                    means, stds = [self.image_mean] * len(images), [self.image_std] * len(images)
            return resize_normalize_batch(images, target_sizes, float(self.max_size), means, stds), targets
        # anything else (CPU tensors, other dtypes) runs the reference's own torch sequence: this is the transform of the
        # detector's input, not the blur path, and must keep working for inputs the CUDA pass does not take
        for i in range(len(images)):
            image = images[i]
            if image.dim() != 3:
                raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W], got {}".format(image.shape))
            if self.normalize_images:
                if newMeans is not None:
                    image = self.normalize(image, newMeans[i, :], newSTDs[i, :])
                else:
                    image = self.normalize(image, self.image_mean, self.image_std)
            if self.training:                        # torch_choice (:141-149)
                size = float(self.min_size[int(torch.empty(1).uniform_(0., float(len(self.min_size))).item())])
            else:
                size = float(self.min_size[-1])
            scale = resize_scale(image.shape[-2], image.shape[-1], size, self.max_size)
            if scale != 1.0:
                image = torch.nn.functional.interpolate(image[None], scale_factor=scale, mode='bilinear',
                                                        recompute_scale_factor=True, align_corners=False)[0]
            images[i] = image
        sizes = [(int(img.shape[-2]), int(img.shape[-1])) for img in images]
        hp, wp = padded_batch_shape(sizes)
        batch = images[0].new_full((len(images), images[0].shape[0], hp, wp), 0)
        for img, pad_img in zip(images, batch):
            pad_img[:, :img.shape[1], :img.shape[2]].copy_(img)
        return ImageList(batch, sizes), targets
