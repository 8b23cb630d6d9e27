"""fp16 batch of config 2 through (a) the tiled kernel's fused half I/O, (b) torch casts around the fp32 tiled kernel."""
import sys
import torch
sys.path.insert(0, ".")
import bench
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops

dev = torch.device("cuda")
spec = bench.workload_spec("cfg2h", None)
traj, fr = bench.make_trajectories(spec, 0)
psfs = ops.rasterize_psfs(traj, fr, dev, dtype=torch.float16)
ts = ops.compact_taps(psfs, normalize=True)
gen = torch.Generator().manual_seed(0)
batches = [torch.rand((8, 3, 800, 1333), generator=gen).half().to(dev) for _ in range(3)]


def timed(fn, n=30):
    for k in range(5):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(n):
        fn(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


idx = list(range(8))
fused = timed(lambda k: bf.blur_batch([batches[k % 3][i] for i in range(8)], ts, idx))
casts = timed(lambda k: bf.blur_batch([batches[k % 3][i] for i in range(8)], ts, idx, clamp=[False] * 8))
a = bf.blur_batch([batches[0][i] for i in range(8)], ts, idx)
b = bf.blur_batch([batches[0][i] for i in range(8)], ts, idx, clamp=[False] * 8)
print("fused half i/o: %.1f us per batch; torch casts around the fp32 kernel: %.1f us; identical: %s" % (
    fused, casts, all(torch.equal(x, y) for x, y in zip(a, b))))
