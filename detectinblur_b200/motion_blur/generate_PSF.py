"""Mirror of ``motion_blur/generate_PSF.py``: the ``PSF`` class, rasterised by the CUDA kernel.

``PSF(canvas, trajectory, fraction).fit()`` / ``.centerPSF()`` / ``.findOffsets()`` / ``.PSFs`` keep the
reference's contract (generate_PSF.py:10-150); the 2000-iteration Python splat loop (:31-77) and the Python centroid
loop (:106-123) are replaced by ``dib_rasterize_psf`` (one CTA per PSF, bit-identical fp64 results).  ``device``
selects the GPU; there is no CPU path -- the class raises without CUDA, and DataLoader workers (which must not touch
CUDA) should hand the trajectory to the main process instead (see transforms.BlurImage(psf_backend="defer")).
"""
import numpy as np
import torch

from .. import psf_ops
from .generate_trajectory import Trajectory


class PSF(object):
    def __init__(self, canvas=None, trajectory=None, fraction=None, path_to_save=None, device=None):
        self.canvas = (canvas, canvas)
        if trajectory is None:
            self.trajectory_obj = Trajectory(canvas=canvas, expl=0.005).fit()
            self.trajectory = self.trajectory_obj.x
        else:
            self.trajectory = trajectory.x
        self.fraction = [1 / 100, 1 / 10, 1 / 2, 1] if fraction is None else fraction
        self.path_to_save = path_to_save
        self.PSFnumber = len(self.fraction)
        self.iters = len(self.trajectory)
        self.PSFs = []
        self.device = torch.device(device) if device is not None else torch.device("cuda")

    def fit(self, show=False, save=False):
        """generate_PSF.py:31-83.  With several fractions the reference accumulates consecutive exposure slices into one
        running canvas (PSF j = slices 0..j); only single-fraction lists are used on the blur path (transforms.py:320)."""
        if self.PSFnumber != 1:
            raise NotImplementedError("multi-fraction PSF lists are not on the blur path (transforms.py:320 passes one fraction)")
        if show or save:
            raise NotImplementedError("plotting is out of scope (matplotlib-only code in the reference)")
        psf = psf_ops.rasterize_psfs(self.trajectory[None], [self.fraction[0]], self.device, canvas=self.canvas[0],
                                     center=False, out_side=self.canvas[0], dtype=torch.float64)
        self.PSFs.append(psf[0].cpu().numpy())
        return self.PSFs

    def centerPSF(self):
        """generate_PSF.py:106-123: roll the PSF so that its weighted centroid sits at the canvas centre."""
        psf = psf_ops.rasterize_psfs(self.trajectory[None], [self.fraction[0]], self.device, canvas=self.canvas[0],
                                     center=True, out_side=self.canvas[0], dtype=torch.float64)
        self.PSFs[0] = psf[0].cpu().numpy()

    def findOffsets(self):
        """generate_PSF.py:125-147: [left, top, right, bottom] extents of the support around canvas/2 - 1."""
        ys, xs = np.nonzero(self.PSFs[0] > 0)
        cx, cy = self.canvas[0] / 2 - 1, self.canvas[1] / 2 - 1
        ox, oy = xs - cx, ys - cy
        right = max(0, ox.max()) if len(ox) else 0
        left = max(0, (-ox[ox <= 0]).max()) if np.any(ox <= 0) else 0
        bottom = max(0, oy.max()) if len(oy) else 0
        top = max(0, (-oy[oy <= 0]).max()) if np.any(oy <= 0) else 0
        return [left, top, right, bottom]
