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

# kernel-only time of the compaction (no meta read-back): the C entry point called directly between two events
import ctypes
from detectinblur_b200 import _lib
for name in ("cfg2", "cfg3"):
    spec = workload_spec(name, None)
    traj, fr = make_trajectories(spec, 0)
    psfs = ops.rasterize_psfs(traj, fr, "cuda", dtype=torch.float32).contiguous()
    n = len(fr)
    lay = _lib.tapset_layout(n, 1024)
    buf = torch.empty(lay.total_bytes, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        _lib.check(_lib.lib.dib_compact_taps(ctypes.c_void_p(psfs.data_ptr()), 0, n, 128, 128 * 128, 1, ctypes.c_void_p(buf.data_ptr()), 1024, st))
    print(name, "dib_compact_taps kernel only: %.1f us per batch of %d" % (timed(call, 50), n))
