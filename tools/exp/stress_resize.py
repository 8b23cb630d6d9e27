"""Randomised stress of the fused normalize + bilinear resize + zero-padded batch kernel (dib_resize_normalize_batch) against the
reference's own sequence of torch calls (models/net_transforms.py:112-133, 151-175, 218-249) on the same device.
    python tools/exp/stress_resize.py [seconds] [seed]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from detectinblur_b200 import net_transforms as nt


def run(budget=20.0, seed=0):
    rng = np.random.default_rng(seed)
    dev = torch.device("cuda")
    t0 = time.time()
    n_cases, worst = 0, 0.0
    while time.time() - t0 < budget:
        n = int(rng.integers(1, 9))
        min_size, max_size = [(800.0, 1333.0), (600.0, 1000.0), (320.0, 480.0), (1024.0, 1400.0)][int(rng.integers(0, 4))]
        imgs = [torch.rand((3, int(rng.integers(40, 900)), int(rng.integers(40, 1400))), device=dev) for _ in range(n)]
        means = [list(rng.random(3) * 0.5 + 0.2) for _ in range(n)]
        stds = [list(rng.random(3) * 0.3 + 0.1) for _ in range(n)]
        il = nt.resize_normalize_batch(imgs, min_size, max_size, means, stds)
        outs = []
        for im, m, s in zip(imgs, means, stds):
            x = nt.normalize(im, m, s)
            sc = nt.resize_scale(im.shape[1], im.shape[2], min_size, max_size)
            outs.append(torch.nn.functional.interpolate(x[None], scale_factor=sc, mode="bilinear", recompute_scale_factor=True,
                                                        align_corners=False)[0])
        hp, wp = nt.padded_batch_shape([(int(o.shape[1]), int(o.shape[2])) for o in outs])
        want = outs[0].new_full((n, 3, hp, wp), 0)
        for o, pad in zip(outs, want):
            pad[:, :o.shape[1], :o.shape[2]].copy_(o)
        if tuple(il.tensors.shape) != tuple(want.shape) or [tuple(s) for s in il.image_sizes] != [(int(o.shape[1]), int(o.shape[2])) for o in outs]:
            print("SHAPE MISMATCH", tuple(il.tensors.shape), tuple(want.shape), il.image_sizes)
            raise AssertionError("batch geometry differs from the reference transform")
        err = float((il.tensors - want).abs().max())
        worst = max(worst, err)
        if not err <= 2e-5:
            print("MISMATCH", dict(case=n_cases, n=n, sizes=[tuple(i.shape) for i in imgs], min_size=min_size, max_size=max_size, err=err))
            raise AssertionError("resized batch differs from the reference transform")
        n_cases += 1
    msg = "ok: %d batches in %.0f s, worst difference %.3g" % (n_cases, time.time() - t0, worst)
    print(msg)
    return msg


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 20.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
