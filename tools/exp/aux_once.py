"""Launch every auxiliary kernel a couple of times (for an ncu capture): trajectories, rasteriser, tap compaction,
packed-bank unpack, fused resize pass."""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, ".")
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import net_transforms as nt
from detectinblur_b200 import psf_bank

dev = torch.device("cuda")
params = np.array([0.005, 0.001, 0.00005] * 86)[:256]
fracs = np.array([1 / 25, 1 / 10, 1 / 5, 1 / 2, 1] * 52)[:256]
for _ in range(2):
    traj = ops.generate_trajectories(256, params, 1, dev)
    psfs16 = ops.rasterize_psfs(traj, fracs, dev, canvas=256, center=True, out_side=256, dtype=torch.float16)
    psfs = ops.rasterize_psfs(traj[:16], fracs[:16], dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
    ts = ops.compact_taps(psfs, normalize=True)
with tempfile.TemporaryDirectory() as d:
    psf_bank.write_pack(os.path.join(d, "P1E0.dibpack"), 0, list(psfs16[:32].cpu().numpy()))
    bank = psf_bank.PackedPsfBank(d)
    for _ in range(2):
        up = bank.upload([(1, 0, k) for k in range(32)], dev)
gen = torch.Generator().manual_seed(0)
sizes = [(480, 640), (427, 640), (640, 480), (640, 427), (375, 500), (500, 375), (333, 500), (480, 640)]
imgs = [torch.rand((3,) + s, generator=gen).to(dev) for s in sizes]
for _ in range(2):
    il = nt.resize_normalize_batch(imgs, 800.0, 1333.0, [nt.CANONICAL_MEAN] * 8, [nt.CANONICAL_STD] * 8)
torch.cuda.synchronize()
print("ok", tuple(il.tensors.shape), ts.counts[:4], tuple(up.shape))
