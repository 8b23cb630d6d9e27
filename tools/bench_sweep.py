"""Secondary bench, BASELINE config 4: the on-the-fly evaluation sweep.

P in {0.005, 0.001, 0.00005} x E in {1/25, 1/10, 1/5, 1/2, 1} (evaluate.py:299-300); one seeded host trajectory per cell
(``Trajectory.fit().fit()``, excluded from timing: it stays host code); images of COCO-val-like sizes
cycled over the cells.  Timed per step, on the device with CUDA events and by wall clock (the path is launch-bound, so
host time is part of the story): rasterise the PSFs -> compact taps (normalise) -> blur.  Two shapes of the same work:

  * batch 1, as the reference's eval loop does it (one image per iteration): 3 launches per image;
  * the 15 cells as one batch: 3 launches (+1 when a PSF needs the exact-order kernel) for 15 images.

Also reports the rasteriser alone (PSFs/s at batch 256).  Prints one JSON object; run on the GPU box:

    python tools/bench_sweep.py [--repeats 50] > gpurun_out/sweep.json
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import detectinblur_b200.blur_functions as bf  # noqa: E402
import detectinblur_b200.psf_ops as ops  # noqa: E402
from detectinblur_b200.motion_blur.generate_trajectory import Trajectory  # noqa: E402

PARAMS = [0.005, 0.001, 0.00005]
EXPOSURES = [1 / 25, 1 / 10, 1 / 5, 1 / 2, 1]
SIZES = [(480, 640), (427, 640), (640, 480), (640, 427), (375, 500), (500, 375), (333, 500)]


def timed(fn, repeats):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(repeats):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / repeats * 1e3, (time.perf_counter() - t0) / repeats * 1e6     # device us, wall us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeats", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda")
    np.random.seed(1337)
    random.seed(1337)
    cells = [(p, e) for p in PARAMS for e in EXPOSURES]
    traj = np.stack([Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x for p, _ in cells])
    frac = np.array([e for _, e in cells])
    gen = torch.Generator(device="cpu").manual_seed(1337)
    images = [torch.rand((3,) + SIZES[k % len(SIZES)], generator=gen).to(dev) for k in range(len(cells))]
    n = len(cells)
    l0 = bf.launch_count()

    def one(k):
        psf = ops.rasterize_psfs(traj[k:k + 1], frac[k:k + 1], dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
        ts = ops.compact_taps(psf, normalize=True)
        return bf.blur_batch([images[k]], ts, [0])[0], ts

    per_cell = []
    for k, (p, e) in enumerate(cells):
        dev_us, wall_us = timed(lambda: one(k), args.repeats)
        _, ts = one(k)
        from detectinblur_b200 import _lib
        psf_k = ops.rasterize_psfs(traj[k:k + 1], frac[k:k + 1], dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
        r_us, _ = timed(lambda: ops.rasterize_psfs(traj[k:k + 1], frac[k:k + 1], dev, canvas=256, center=True, out_side=128,
                                                   dtype=torch.float32), args.repeats)
        c_us, _ = timed(lambda: ops.compact_taps(psf_k, normalize=True), args.repeats)
        plan = bf.prepare_blur([images[k]], ts, [0])
        b_us, _ = timed(plan.run, args.repeats)
        per_cell.append({"param": p, "exposure": e, "size": list(SIZES[k % len(SIZES)]), "taps": ts.counts[0],
                         "kernel": "generic" if ts.meta[0].flags & _lib.META_NO_PROGRAM else "tiled",
                         "device_us": round(dev_us, 1), "wall_us": round(wall_us, 1), "rasterize_us": round(r_us, 1),
                         "compact_us": round(c_us, 1), "blur_us": round(b_us, 1)})

    # ---- batch 1 without a host synchronisation: device-resident trajectories, compact_taps(sync=False) (the launch is
    #      planned on the device), results into preallocated buffers; issued directly and replayed from a CUDA graph
    d_traj = [torch.from_numpy(np.ascontiguousarray(traj[k:k + 1])).to(dev) for k in range(n)]
    d_frac = [torch.tensor([frac[k]], dtype=torch.float64, device=dev) for k in range(n)]
    d_outs = [torch.empty((3, images[k].shape[1], (images[k].shape[2] + 3) // 4 * 4), device=dev)[:, :, :images[k].shape[2]] for k in range(n)]

    def one_async(k):
        psf = ops.rasterize_psfs(d_traj[k], d_frac[k], dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
        ts = ops.compact_taps(psf, normalize=True, max_taps=4096, sync=False)
        bf.blur_batch([images[k]], ts, [0], outs=[d_outs[k]])
        return ts

    async_cells = []
    for k in range(n):
        a_dev, a_wall = timed(lambda: one_async(k), args.repeats)
        g_dev = g_wall = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                keep = one_async(k)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                keep = one_async(k)
            g_dev, g_wall = timed(graph.replay, args.repeats)
            del keep
        except Exception as exc:
            sys.stderr.write("graph capture failed for cell %d: %s\n" % (k, str(exc).splitlines()[0]))
        async_cells.append({"direct_device_us": round(a_dev, 1), "direct_wall_us": round(a_wall, 1),
                            "graph_device_us": None if g_dev is None else round(g_dev, 1),
                            "graph_wall_us": None if g_wall is None else round(g_wall, 1)})

    def batched():
        psf = ops.rasterize_psfs(traj, frac, dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
        ts = ops.compact_taps(psf, normalize=True)
        return bf.blur_batch(images, ts, list(range(n)))

    bdev, bwall = timed(batched, args.repeats)
    big = np.concatenate([traj] * 18)[:256]
    bigf = np.concatenate([frac] * 18)[:256]
    t_traj = None

    def raster():
        return ops.rasterize_psfs(big, bigf, dev, canvas=256, center=True, out_side=128, dtype=torch.float16)

    rdev, rwall = timed(raster, 10)
    # the input transform after the blur: normalize + resize to 800 / 1333 + zero-padded batch, 8 COCO-size images
    from detectinblur_b200 import net_transforms as nt
    timgs = [images[k] for k in range(8)]
    means = [nt.CANONICAL_MEAN] * 8
    stds = [nt.CANONICAL_STD] * 8
    fused_us, _ = timed(lambda: nt.resize_normalize_batch(timgs, 800.0, 1333.0, means, stds), args.repeats)
    il = nt.resize_normalize_batch(timgs, 800.0, 1333.0, means, stds)

    def torch_ops():       # the reference's own sequence of torch calls (net_transforms.py:112-133) on the same device
        outs = []
        for im in timgs:
            x = nt.normalize(im, nt.CANONICAL_MEAN, nt.CANONICAL_STD)
            sc = nt.resize_scale(im.shape[1], im.shape[2], 800.0, 1333.0)
            outs.append(torch.nn.functional.interpolate(x[None], scale_factor=sc, mode="bilinear", recompute_scale_factor=True,
                                                        align_corners=False)[0])
        hp, wp = nt.padded_batch_shape([(int(o.shape[1]), int(o.shape[2])) for o in outs])
        batch = outs[0].new_full((len(outs), 3, hp, wp), 0)
        for o, pad in zip(outs, batch):
            pad[:, :o.shape[1], :o.shape[2]].copy_(o)
        return batch

    torch_us, _ = timed(torch_ops, args.repeats)
    tdiff = float((torch_ops() - il.tensors).abs().max())
    tbytes = sum(im.numel() * 4 for im in timgs) + il.tensors.numel() * 4
    # trajectories on the device (opt-in, statistical parity): the whole on-the-fly chain without host RNG work
    tr_expl = np.array([p for p, _ in cells] * 18)[:256]
    g_us, _ = timed(lambda: ops.generate_trajectories(256, tr_expl, 1, dev), 5)

    def chain():
        t = ops.generate_trajectories(n, [p for p, _ in cells], 1, dev)
        psf = ops.rasterize_psfs(t, frac, dev, canvas=256, center=True, out_side=128, dtype=torch.float32)
        return bf.blur_batch(images, ops.compact_taps(psf, normalize=True), list(range(n)))

    chain_us, _ = timed(chain, 10)
    t0 = time.perf_counter()
    for p, _ in cells:
        Trajectory(canvas=256, max_len=96, expl=p).fit().fit()
    host_traj_us = (time.perf_counter() - t0) / n * 1e6
    b1_dev = sum(c["device_us"] for c in per_cell)
    b1_wall = sum(c["wall_us"] for c in per_cell)
    print(json.dumps({
        "workload": "cfg4: on-the-fly eval sweep, 3 params x 5 exposures, COCO-like sizes, fp32, rasterise + compact + blur",
        "batch1": {"images_per_s_device": n / (b1_dev * 1e-6), "images_per_s_wall": n / (b1_wall * 1e-6),
                   "launches_per_image": 3, "cells": per_cell},
        "batch1_no_host_sync": {
            "what": "rasterise (device trajectory) -> compact_taps(sync=False) -> device-planned blur: 5 launches, no read-back",
            "images_per_s_wall_direct": n / (sum(c["direct_wall_us"] for c in async_cells) * 1e-6),
            "images_per_s_wall_cuda_graph": (n / (sum(c["graph_wall_us"] for c in async_cells) * 1e-6)
                                             if all(c["graph_wall_us"] for c in async_cells) else None),
            "images_per_s_device_cuda_graph": (n / (sum(c["graph_device_us"] for c in async_cells) * 1e-6)
                                               if all(c["graph_device_us"] for c in async_cells) else None),
            "cells": async_cells},
        "batched15": {"device_us": bdev, "wall_us": bwall, "images_per_s_device": n / (bdev * 1e-6),
                      "images_per_s_wall": n / (bwall * 1e-6)},
        "rasterizer_batch256": {"device_us": rdev, "wall_us_incl_h2d": rwall, "psfs_per_s_device": 256 / (rdev * 1e-6),
                                "psfs_per_s_wall": 256 / (rwall * 1e-6)},
        "transform_8_images": {"batch_shape": list(il.tensors.shape), "fused_kernel_us": fused_us, "torch_ops_us": torch_us,
                               "algorithmic_bytes": tbytes, "fused_gbs": tbytes / (fused_us * 1e-6) / 1e9,
                               "max_abs_diff_vs_torch_cuda": tdiff},
        "gpu_trajectories": {"batch256_device_us": g_us, "trajectories_per_s": 256 / (g_us * 1e-6),
                             "chain15_generate_rasterize_compact_blur_us": chain_us,
                             "chain15_images_per_s": n / (chain_us * 1e-6),
                             "host_trajectory_us_each (Trajectory.fit().fit(), this box)": host_traj_us},
        "gpu_launches": bf.launch_count() - l0 + nt.launch_count(),
        "reference_cost_note": "the reference spends ~18 ms (trajectory) + ~45 ms (PSF splat loop) + O(taps) ATen launches per image",
    }))


if __name__ == "__main__":
    main()
