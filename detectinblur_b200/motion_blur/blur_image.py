"""Mirror of ``motion_blur/blur_image.py``: ``BlurImageHandler``, the ``--cpu_blur`` blur, evaluated by the CUDA kernels.

Same constructor, ``blur_image(save, show, oldDeltaPad) -> True``, ``.pilImageResult`` (uint8 PIL) and ``.result``
([float32 HxWx3]) as /root/reference/motion_blur/blur_image.py:23-154.  What the reference computes with three
full-size FFT convolutions is, for a sparse kernel, a short tap sum:

  * the image is edge-padded by half the kernel and min-max normalised to [0, 1]                      (:77-85, :129-130)
  * the kernel is zero-padded to the image and min-max normalised, i.e. divided by its maximum        (:114-128)
  * ``fftconvolve(image, kernel_canvas, 'same')`` is the zero-boundary convolution of the padded image about the
    kernel origin (cy, cx) = ((yN-1)//2 - top, (xN-1)//2 - left): (63, 63) for even padded sizes, 64 rows when the padded
    height is odd (the reference pads ``ceil`` on the left but ``floor`` on top, :119-123)             (:131-133)
  * the result is min-max stretched over the whole padded frame, unpadded, and truncated to uint8      (:134-147)

The tap sum runs on ``dib_blur_batch`` in zero-padding mode over the padded frame (its dark rim takes part in the
min-max stretch, so it has to be computed); an origin other than (63, 63) is expressed by extending the frame with
zero rows / columns, which a zero-boundary convolution does not see.  The min-max reductions and the final
scale / crop / uint8 conversion are torch calls on the device.  Images smaller than the kernel keep the reference's
host-side PIL bicubic upscale before and OpenCV Lanczos resize after (:56-69, :142-143).

There is no CPU path: the handler raises without CUDA.  Results agree with the reference to <= 1e-5 on ``result[0]`` and
<= 1 level on the uint8 image (tests/test_gpu_transforms.py against tests/golden/fourier_cases.npz).
"""
import math
import os

import numpy as np
import torch

from .. import _lib
from .. import blur_functions
from .. import psf_ops


def kernel_origin(psf_shape, yN, xN, oldDeltaPad=False):
    """(cy, cx) such that output (i, j) reads the padded image at (i + cy - y, j + cx - x) for kernel element (y, x)."""
    key, kex = psf_shape
    dY, dX = yN - key, xN - kex
    if oldDeltaPad:                                    # :116-117 pads dX // 2 on every side
        pad = dX // 2
        return (key + 2 * pad - 1) // 2 - pad, (kex + 2 * pad - 1) // 2 - pad
    return (yN - 1) // 2 - dY // 2, (xN - 1) // 2 - math.ceil(dX / 2)


class BlurImageHandler(object):

    def __init__(self, image_path, PSFs=None, pillowImage=None, part=None, path__to_save=None, buffPadImage=True, device=None):
        from PIL import Image
        self.path_to_save = path__to_save
        if PSFs is None:
            # the reference builds PSF(canvas=self.shape[0]) here before self.shape exists and fails (:34-39)
            raise AttributeError("'BlurImageHandler' object has no attribute 'shape' (pass PSFs=[...])")
        self.PSFs = PSFs
        if pillowImage is None:
            if image_path is not None and os.path.isfile(image_path):
                self.image_path = image_path
                self.original = Image.open(self.image_path)
            else:
                raise Exception('Not correct path to image.')
        else:
            self.original = pillowImage
            self.originalPillowImage = self.original
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        if not torch.cuda.is_available():
            raise RuntimeError("BlurImageHandler runs on the CUDA kernels of detectinblur_b200; no CUDA device is available")

        # resize before the array conversion (:55-69); PIL .size is (W, H) and the reference's names are swapped
        self.originalSize = self.original.size
        yN, xN = self.original.size
        key, kex = self.PSFs[0].shape
        if yN - key < 0 or xN - kex < 0:
            ratioY, ratioX = key / yN, kex / xN
            r = ratioX if ratioX > ratioY else ratioY
            self.original = self.original.resize((math.ceil(r * xN), math.ceil(r * yN)), Image.BICUBIC)
        else:
            self.originalSize = None
        self.original = np.array(self.original)

        self.buffPadImage = buffPadImage
        if buffPadImage:                                                          # :77-85
            paddingR = round(self.PSFs[0].shape[0] / 2)
            paddingC = round(self.PSFs[0].shape[1] / 2)
            if len(self.original.shape) > 2:
                padding = ((paddingR, paddingR), (paddingC, paddingC), (0, 0))
            else:
                padding = ((paddingR, paddingR), (paddingC, paddingC))
            self.original = np.pad(self.original, pad_width=padding, mode='edge')
        if len(self.original.shape) < 3:                                          # :91-97
            self.original = np.stack([self.original] * 3, axis=2)
        self.shape = self.original.shape
        self.part = part
        self.result = []
        self.pilImageResult = None

    def blur_image(self, save=False, show=False, oldDeltaPad=False):
        from PIL import Image
        psf = self.PSFs if self.part is None else [self.PSFs[self.part]]
        psf = np.asarray(psf[0])
        yN, xN, channel = self.shape
        key, kex = self.PSFs[0].shape
        if yN - key < 0 or xN - kex < 0:
            raise ValueError("index can't contain negative values")                # what np.pad raises in the reference (:123)
        if key > 129 or key != kex:
            raise NotImplementedError("BlurImageHandler on the CUDA path takes square kernels up to 129 x 129")
        dev = self.device

        # kernel: min-max normalisation over the zero-padded canvas (:128).  With zeros present (a sparse kernel, or any
        # padding) the minimum is 0 and the map is a division by the maximum.
        pk = torch.as_tensor(psf, dtype=torch.float32, device=dev)
        kmin, kmax = (float(v) for v in torch.aminmax(pk))
        if (yN > key or xN > kex) and kmin > 0.0:
            kmin = 0.0
        if kmin != 0.0:
            raise NotImplementedError("kernels with negative or no zero entries make the normalised canvas dense")
        d = kmax - kmin
        pk = pk * (1.0 / d if d > np.finfo(np.float64).eps else 0.0)

        # image: uint8 HWC -> CHW float32, min-max normalised (:129-130)
        img = torch.from_numpy(np.ascontiguousarray(self.original)).to(dev)
        img = img.permute(2, 0, 1).to(torch.float32)
        imin, imax = (float(v) for v in torch.aminmax(img))
        d = imax - imin
        scale = 1.0 / d if d > np.finfo(np.float64).eps else 0.0
        img = img * scale + (0.0 - imin * scale)

        # zero-boundary convolution about (cy, cx): the kernels' origin is (63, 63); a different origin is a frame shifted
        # by (sy, sx), realised with zero rows / columns that the zero-padding mode cannot tell from the outside
        cy, cx = kernel_origin((key, kex), yN, xN, oldDeltaPad)
        sy, sx = cy - 63, cx - 63
        if sy or sx:
            frame = torch.zeros((channel, yN + abs(sy), xN + abs(sx)), dtype=torch.float32, device=dev)
            frame[:, max(-sy, 0):max(-sy, 0) + yN, max(-sx, 0):max(-sx, 0) + xN] = img
        else:
            frame = img.contiguous()
        tapset = psf_ops.compact_taps(pk, normalize=False)
        out = blur_functions.blur_batch([frame], tapset, [0], pad_mode=_lib.PAD_ZERO128)[0]
        out = out[:, max(sy, 0):max(sy, 0) + yN, max(sx, 0):max(sx, 0) + xN]

        # min-max stretch over the padded frame (:134), unpad (:137-140)
        omin, omax = (float(v) for v in torch.aminmax(out))
        d = omax - omin
        scale = 1.0 / d if d > np.finfo(np.float64).eps else 0.0
        blured = out * scale + (0.0 - omin * scale)
        if self.buffPadImage:
            paddingR = round(self.PSFs[0].shape[0] / 2)
            paddingC = round(self.PSFs[0].shape[1] / 2)
            blured = blured[:, paddingR:blured.shape[1] - paddingR, paddingC:blured.shape[2] - paddingC]
        blured = blured.permute(1, 2, 0).contiguous()

        if self.originalSize is not None:                                          # :142-143, host side as in the reference
            import cv2
            host = cv2.resize(blured.cpu().numpy(), self.originalSize, interpolation=cv2.INTER_LANCZOS4)
            self.result = [np.abs(host)]
            self.pilImageResult = Image.fromarray((host * 255).astype(np.uint8))
        else:
            u8 = (blured * 255).to(torch.uint8)                                    # truncation, like astype(np.uint8) on [0, 1]
            self.result = [blured.abs().cpu().numpy()]
            self.pilImageResult = Image.fromarray(u8.cpu().numpy())
        if show or save:
            self.plot_canvas(show, save)
        return True

    def plot_canvas(self, show, save):
        """:156-163: write result[0] next to ``path__to_save`` under the input's file name."""
        if len(self.result) == 0:
            raise Exception('Please run blur_image() method first.')
        if self.path_to_save is None:
            raise Exception('Please create Trajectory instance with path_to_save')
        import cv2
        toSave = cv2.cvtColor(self.result[0], cv2.COLOR_RGB2BGR)
        cv2.imwrite(os.path.join(self.path_to_save, self.image_path.split('/')[-1]), toSave * 255)
