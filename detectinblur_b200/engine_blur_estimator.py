"""Mirror of the blur helpers duplicated in the reference's ``engine_blur_estimator.py`` (:27-79).

The estimator's copy of ``manual_blur`` is the 128-px branch of models/blur_functions.py with an optional
``resize_images`` step: bilinear resize to a short side of 800 before the blur (portrait images are transposed
first), then -- as written upstream (:63-68) -- the padded accumulator is cropped with the ORIGINAL height / width and
"resized" to that same size.  For images no larger than 800 px that crop is the top-left original-sized window of the
blurred, resized (and, for portraits, still transposed) image; this mirror reproduces exactly that.  Larger images make
the reference read its padding accumulator, which the kernel never materialises: they raise NotImplementedError.
"""
import torch

from . import blur_functions
from . import psf_ops


def manual_blur(image_GPU, psf_GPU, resize_images=False, exact=None):
    """engine_blur_estimator.py:27-70."""
    psf_ops._require_cuda(image_GPU, "image_GPU")
    image_height, image_width = int(image_GPU.shape[1]), int(image_GPU.shape[2])
    work = image_GPU
    if resize_images:
        if image_height > 800 or image_width > 800:
            raise NotImplementedError("resize_images on images larger than 800 px reads the reference's padding accumulator "
                                      "(engine_blur_estimator.py:63); not reproduced")
        x = image_GPU.unsqueeze(0)
        if image_height > image_width:
            x = x.permute(0, 1, 3, 2)                                        # :34-35 (never undone upstream)
            new_height, new_width = 800, int(800 * image_height / image_width)
        else:
            new_height, new_width = 800, int(800 * image_width / image_height)
        work = torch.nn.functional.interpolate(x, size=(new_height, new_width), mode="bilinear")[0].contiguous()   # :42
    if psf_GPU.shape[0] > 129:
        raise ValueError("the estimator's blur only has the 128-px branch (engine_blur_estimator.py:45-49)")
    blurred = blur_functions.manual_blur(work, psf_GPU, exact=exact)
    if blurred.dim() == 2:
        blurred = blurred.unsqueeze(0)
    if resize_images:
        window = blurred[:, :image_height, :image_width].unsqueeze(0)           # :61 crop with the original extent
        blurred = torch.nn.functional.interpolate(window, size=(image_height, image_width), mode="bilinear")[0]   # :68
    return blurred.squeeze()


def blur_image_list(images_GPU, blur_dicts, psfs_GPU, resize_images=False, exact=None):
    """engine_blur_estimator.py:72-79: in place over the list, PSFs normalised by their own sum."""
    if not resize_images:
        return blur_functions.blur_image_list(images_GPU, blur_dicts, psfs_GPU, exact=exact)
    for k, (img, bd, psf) in enumerate(zip(images_GPU, blur_dicts, psfs_GPU)):
        if not bd["blurring"]:
            continue
        images_GPU[k] = manual_blur(img, psf / psf.sum(), resize_images=True, exact=exact)
    return None
