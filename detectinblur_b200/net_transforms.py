"""Mirror of the input-transform step of the reference's ``models/net_transforms.py`` that follows the blur.

``GeneralizedRCNNTransform.forward`` (net_transforms.py:82-133) normalises every image with its own row of
``newMeans`` / ``newSTDs`` (``normalize`` :135-139), resizes (:151-175) and packs the list into one zero-padded batch
whose sides are multiples of 32 (``batch_images`` :218-249).  Here the normalisation is the epilogue of the blur
kernel and the result lands directly inside the padded batch tensor (``fused_blur_normalize``), so blurred pixels
are written once; a model built with ``GeneralizedRCNNTransform(..., normalize_images=False)`` -- the hook the
reference already has (:70-80, :112-118) -- then consumes the batch unchanged.  Resizing stays torch code and only
runs when an image is not already at the requested scale (identity at 800x1333).
"""
import math

import torch

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


def fused_blur_normalize(images_GPU, blur_dicts, psfs_GPU, newMeans=None, newSTDs=None, image_mean=None, image_std=None,
                         size_divisible=32, exact=None):
    """blur_image_list + normalize + batch_images in one pass over the pixels.

    images_GPU   list of [3, H, W] CUDA tensors (float32), as handed to blur_image_list
    blur_dicts   per image; entries with a falsy "blurring" are only normalised
    psfs_GPU     per image dense PSF (ignored where not blurring)
    newMeans / newSTDs   optional [N, 3] per-image statistics (utils.get_norm_params, utils.py:219-273); the
                 canonical ImageNet statistics otherwise
    Returns ImageList(tensors=[N, 3, Hp, Wp] zero padded, image_sizes=[(H, W), ...]).
    """
    n = len(images_GPU)
    if n == 0:
        raise ValueError("empty image list")
    image_mean = CANONICAL_MEAN if image_mean is None else image_mean
    image_std = CANONICAL_STD if image_std is None else image_std
    dev, dtype = images_GPU[0].device, images_GPU[0].dtype
    sizes = [(int(im.shape[1]), int(im.shape[2])) for im in images_GPU]
    hp, wp = padded_batch_shape(sizes, size_divisible)
    batch = torch.zeros((n, int(images_GPU[0].shape[0]), hp, wp), dtype=dtype, device=dev)
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
        stack = torch.stack([psfs_GPU[k].to(dtype) for k in blurred])
        tapset = psf_ops.compact_taps(stack, normalize=True)
        for j, k in enumerate(blurred):
            idx[k] = j
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
        self.crop_images = crop_images
        self.training = training
        self.normalize_images = normalize_images

    def normalize(self, image, image_mean, image_std):
        return normalize(image, image_mean, image_std)

    def forward(self, images, targets=None, newMeans=None, newSTDs=None):
        if isinstance(images, ImageList):
            return images, targets                   # already normalised and batched by fused_blur_normalize
        images = [img for img in images]
        for i in range(len(images)):
            image = images[i]
            if image.dim() != 3:
                raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W], got {}".format(image.shape))
            if self.normalize_images:
                if newMeans is not None:
                    image = self.normalize(image, newMeans[i, :], newSTDs[i, :])
                else:
                    image = self.normalize(image, self.image_mean, self.image_std)
            scale = resize_scale(image.shape[-2], image.shape[-1], self.min_size[-1], self.max_size)
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
