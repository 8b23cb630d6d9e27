"""Batch-1 on-the-fly chain (rasterise -> compact(sync=False) -> device-planned blur), a few iterations per chosen sweep cell.
Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel latency of the chain, or alone for event timings."""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200.motion_blur import Trajectory

PARAMS = [0.005, 0.001, 0.00005]
EXPOSURES = [1 / 25, 1 / 10, 1 / 5, 1 / 2, 1]


def main():
    dev = torch.device("cuda")
    np.random.seed(1337)
    random.seed(1337)
    which = [int(a) for a in sys.argv[1:]] or [0, 4, 14]
    cells = [(p, e) for p in PARAMS for e in EXPOSURES]
    traj = np.stack([Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x for p, _ in cells])
    img = torch.rand((3, 480, 640), device=dev)
    out = torch.empty((3, 480, 640), device=dev)
    for k in which:
        d_traj = torch.from_numpy(np.ascontiguousarray(traj[k:k + 1])).to(dev)
        d_frac = torch.tensor([cells[k][1]], dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(4):
            torch.cuda.synchronize()
            e0.record()
            psf = ops.rasterize_psfs(d_traj, d_frac, dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
            ts = ops.compact_taps(psf, normalize=True, max_taps=4096, sync=False)
            bf.blur_batch([img], ts, [0], outs=[out])
            e1.record()
            torch.cuda.synchronize()
        print("cell", k, cells[k], "chain us (direct, last)", round(e0.elapsed_time(e1) * 1e3, 1))


if __name__ == "__main__":
    main()
