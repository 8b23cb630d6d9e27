"""cProfile of the host side of the batch-1 chain (rasterise -> compact(sync=False) -> device-planned blur), issued directly."""
import cProfile, os, pstats, random, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200.motion_blur import Trajectory

dev = torch.device("cuda")
np.random.seed(1337); random.seed(1337)
traj = Trajectory(canvas=256, max_len=96, expl=0.005).fit().fit().x
d_traj = torch.from_numpy(np.ascontiguousarray(traj[None])).to(dev)
d_frac = torch.tensor([0.1], dtype=torch.float64, device=dev)
img = torch.rand((3, 480, 640), device=dev)
out = torch.empty((3, 480, 640), device=dev)

def one():
    psf = ops.rasterize_psfs(d_traj, d_frac, dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
    ts = ops.compact_taps(psf, normalize=True, max_taps=4096, sync=False)
    bf.blur_batch([img], ts, [0], outs=[out])

for _ in range(200): one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): one()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue %.1f us per image (wall incl. GPU %.1f)" % ((t1 - t0) / 2000 * 1e6, (time.perf_counter() - t0) / 2000 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): one()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
