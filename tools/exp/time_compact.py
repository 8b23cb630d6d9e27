"""Device time of dib_compact_taps alone (no read-back) for 1 / 8 / 16 PSFs of configs 2 and 3."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, bench
import detectinblur_b200.psf_ops as ops
dev = torch.device("cuda")
for name in ("cfg2", "cfg3"):
    spec = bench.workload_spec(name, None)
    traj, fr = bench.make_trajectories(spec, seed=0)
    psfs = ops.rasterize_psfs(traj, fr, dev, dtype=torch.float16).float()
    for n in (1, 8, len(psfs)):
        p = psfs[:n].contiguous()
        for _ in range(5):
            ops.compact_taps(p, normalize=True, sync=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            ops.compact_taps(p, normalize=True, sync=False)
        e1.record()
        torch.cuda.synchronize()
        direct = e0.elapsed_time(e1) / 50 * 1e3
        # the same 20 calls replayed from a CUDA graph: device time without the host's launch path
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep = [ops.compact_taps(p, normalize=True, sync=False) for _ in range(20)]
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(name, "n_psfs", n, "compact_taps %.1f us issued directly, %.1f us from a graph" % (direct, e0.elapsed_time(e1) / 200 * 1e3))
    t = torch.from_numpy(traj[:1]).to(dev)
    f = torch.tensor(fr[:1], dtype=torch.float64, device=dev)
    for _ in range(5):
        ops.rasterize_psfs(t, f, dev, dtype=torch.float32)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        ops.rasterize_psfs(t, f, dev, dtype=torch.float32)
    e1.record()
    torch.cuda.synchronize()
    print(name, "rasterize 1 PSF (fraction %.3f) %.1f us" % (fr[0], e0.elapsed_time(e1) / 50 * 1e3))
