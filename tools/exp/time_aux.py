"""Device time of the auxiliary kernels (tap compaction, PSF rasterisation) with CUDA events."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import detectinblur_b200.psf_ops as ops
from bench import make_trajectories, workload_spec


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name in ("cfg2", "cfg3"):
    spec = workload_spec(name, None)
    traj, fr = make_trajectories(spec, 0)
    psfs = ops.rasterize_psfs(traj, fr, "cuda", dtype=torch.float32)
    print(name, "compact_taps (incl. meta D2H + sync): %.1f us per batch of %d" % (timed(lambda: ops.compact_taps(psfs, True)), len(fr)))
    big = np.concatenate([traj] * 32)[:256]
    frb = np.concatenate([fr] * 32)[:256]
    t = timed(lambda: ops.rasterize_psfs(big, frb, "cuda", dtype=torch.float16), 5)
    print(name, "rasterize 256 PSFs (incl. H2D of trajectories): %.0f us -> %.0f PSFs/s" % (t, 256 / t * 1e6))
