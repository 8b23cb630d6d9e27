"""Stored-PSF bank writer: mirror of the reference's ``dataset_utils/generate_PSFs.py`` (:16-60).

Same directory layout and file format (``<destination>psfs/P{1..3}E{0..4}/I{index:06d}``, ``np.save`` of a float16
256 x 256 canvas, no extension), same seeding (``1337 * worker_index`` for numpy and python RNG) and the same draw order
(param-major, then exposure, then index; ``Trajectory.fit().fit()`` per PSF), so a worker writes byte-identical files.
The trajectories are drawn on the host; rasterisation, centring and the float16 cast run on the GPU in batches.
"""
import os
import random

import numpy as np
import torch

from . import psf_ops
from .motion_blur.generate_trajectory import Trajectory

PARAMS = [0.005, 0.001, 0.00005]                  # dataset_utils/generate_PSFs.py:31
FRACTIONS = [1 / 18, 1 / 10, 1 / 5, 1 / 2, 1]     # :32


def generate_psf_bank(destination_path, worker_index=0, num_workers=12, total_num_psfs=12000, device="cuda", batch=256):
    """Write this worker's slice of the bank; returns the number of files written."""
    slice_size = int(total_num_psfs / num_workers)
    start_index = slice_size * worker_index
    end_index = start_index + slice_size
    np.random.seed(1337 * worker_index)
    random.seed(1337 * worker_index)
    for p in range(len(PARAMS)):
        for e in range(len(FRACTIONS)):
            os.makedirs(destination_path + "psfs/P" + str(p + 1) + "E" + str(e), exist_ok=True)
    written = 0
    for p, param in enumerate(PARAMS):
        for e, exposure in enumerate(FRACTIONS):
            folder = destination_path + "psfs/P" + str(p + 1) + "E" + str(e)
            for lo in range(start_index, end_index, batch):
                hi = min(lo + batch, end_index)
                traj = np.stack([Trajectory(canvas=256, max_len=96, expl=param).fit().fit().x for _ in range(lo, hi)])
                psfs = psf_ops.rasterize_psfs(traj, [exposure] * (hi - lo), device, canvas=256, center=True, out_side=256,
                                              dtype=torch.float16).cpu().numpy()
                for k, index in enumerate(range(lo, hi)):
                    with open(folder + "/I" + "{:06d}".format(index), "wb") as f:
                        np.save(f, psfs[k])
                    written += 1
    return written


def load_stored_psf(stored_psf_directory, param_index, fraction_index, psf_index):
    """The reader side, transforms.py:301-309: np.load + central 128 crop."""
    path = stored_psf_directory + "/P" + str(param_index) + "E" + str(fraction_index) + "/I" + "{:06d}".format(psf_index)
    with open(path, "rb") as f:
        psf = np.load(f)
    if psf.shape[0] > 128:
        psf = psf[64:128 + 64, 64:128 + 64]
    return psf
