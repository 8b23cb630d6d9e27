"""Randomised parity stress (kernel experiments, not a pytest): random image sizes / channel counts / dtypes / pad modes /
input pitches and PSFs of every sweep cell, tiled kernels (host- and device-planned) against the exact-order kernel.
    python tools/exp/stress_parity.py [seconds] [seed]"""
import os, random, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200 import _lib
from detectinblur_b200.motion_blur import Trajectory


def run(budget=60.0, seed=0):
    rng = random.Random(seed)
    np.random.seed(seed)
    random.seed(seed)
    dev = torch.device("cuda")
    PARAMS = [0.005, 0.001, 0.00005]
    EXPOSURES = [1 / 25, 1 / 10, 1 / 5, 1 / 2, 1]
    # a pool of PSFs: every sweep cell a few times, plus dilated ones
    trajs, fracs = [], []
    for p in PARAMS:
        for e in EXPOSURES:
            for _ in range(2):
                trajs.append(Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x)
                fracs.append(e)
    pool = ops.rasterize_psfs(np.stack(trajs), np.array(fracs), dev, dtype=torch.float16).float()
    t0 = time.time()
    n_cases = worst32 = worst16 = 0
    kinds = {}
    while time.time() - t0 < budget:
        nb = rng.randint(1, 4)
        half = rng.random() < 0.3
        zero = rng.random() < 0.25
        planned = rng.random() < 0.4
        imgs, idx = [], []
        for _ in range(nb):
            C = rng.choice([1, 2, 3, 3, 3])
            H = rng.choice([65, 66, 70, 97, 128, 200, 333, 480, 640, 801])
            W = rng.choice([65, 67, 100, 223, 224, 225, 447, 448, 449, 500, 640, 895, 897, 1333])
            x = torch.rand((C, H, W), device=dev)
            if rng.random() < 0.5:      # pitched / offset input view
                pad = rng.choice([1, 3, 4, 7])
                big = torch.zeros((C, H, W + pad), device=dev)
                big[:, :, :W] = x
                x = big[:, :, :W]
            if half:
                x = x.half() if x.is_contiguous() else x.half()
            imgs.append(x)
            idx.append(rng.randrange(pool.shape[0]))
        psfs = pool[idx]
        if rng.random() < 0.15:         # a dilated PSF (transforms.py:338-342): many more taps
            k = torch.ones((1, 1, 5, 5), device=dev) / 25
            psfs = torch.nn.functional.conv2d(psfs[:, None], k, padding=2)[:, 0]
        psfs = psfs.half().contiguous() if half else psfs.contiguous()
        ts = ops.compact_taps(psfs, normalize=True, **({"sync": False, "max_taps": 4096} if planned else {}))
        pad_mode = _lib.PAD_ZERO128 if zero else None
        # store paths: the wrapper's own row-aligned results, caller-owned dense tensors (unaligned rows when W is odd) or views of a
        # wider buffer; epilogues: none, fused normalize, pre-drawn noise + clamp
        kw = {}
        mode = rng.random()
        if mode < 0.3:
            kw["outs"] = [torch.empty(tuple(im.shape), dtype=im.dtype, device=dev) for im in imgs]
        elif mode < 0.45:
            kw["outs"] = [torch.empty((im.shape[0], im.shape[1], im.shape[2] + rng.choice([1, 2, 5, 11])), dtype=im.dtype, device=dev)[:, :, :im.shape[2]]
                          for im in imgs]
        epi = rng.random()
        if epi < 0.25:
            kw["mean"] = [[0.485, 0.456, 0.406][:im.shape[0]] for im in imgs]
            kw["std"] = [[0.229, 0.224, 0.225][:im.shape[0]] for im in imgs]
        elif epi < 0.4 and not half and "outs" in kw:
            # pre-drawn noise must have the destination's layout: same storage shape, same view
            kw["noise"] = [torch.randn(tuple(o._base.shape) if o._base is not None else tuple(o.shape), device=dev)[:, :, :o.shape[2]] for o in kw["outs"]]
            kw["noise_sd"] = [0.01 * (k + 1) for k in range(nb)]
            kw["clamp"] = [True] * nb
        if os.environ.get("STRESS_VERBOSE"):
            print(dict(case=n_cases, nb=nb, half=half, zero=zero, planned=planned, shapes=[tuple(i.shape) for i in imgs],
                       strides=[i.stride() for i in imgs], psf=idx, kw=sorted(kw)), flush=True)
        got = bf.blur_batch(imgs, ts, list(range(nb)), pad_mode=pad_mode, **kw)
        torch.cuda.synchronize()
        kw_ref = dict(kw)
        if "outs" in kw_ref:
            kw_ref["outs"] = [torch.empty_strided(tuple(o.shape), o.stride(), dtype=o.dtype, device=dev) for o in kw["outs"]]
        # reference: the exact-order kernel in the images' own dtype (for half images that is the reference's half loop, which
        # rounds after every tap: the tiled path's fp32 accumulation may differ from it by the documented 5e-3, more for
        # hundreds of taps)
        ts_ref = ops.compact_taps(psfs, normalize=True)
        want = bf.blur_batch(imgs, ts_ref, list(range(nb)), pad_mode=pad_mode, exact=True, **kw_ref)
        for k in range(nb):
            err = float((got[k].float() - want[k].float()).abs().max())
            tol = 2e-2 if half else 1e-5      # 270 roundings to half in the reference loop against one
            if "mean" in kw:
                tol *= 5.0                     # (x - mean) / std with std ~ 0.225 scales the difference
            if half:
                worst16 = max(worst16, err)
            else:
                worst32 = max(worst32, err)
            if not (err <= tol):
                print("MISMATCH", dict(case=n_cases, k=k, shape=tuple(imgs[k].shape), stride=imgs[k].stride(), half=half, zero=zero, planned=planned,
                                       psf=idx[k], err=err, kw=sorted(kw)))
                raise AssertionError("mismatch (see the line printed above)")
        if not planned:
            for m in ts.meta:
                kinds[m.prog_group_w] = kinds.get(m.prog_group_w, 0) + 1
        n_cases += 1
    msg = ("ok: %d batches in %.0f s, worst fp32 err %.3g, worst fp16 err %.3g, program kinds (group width: count) %s" % (
        n_cases, time.time() - t0, worst32, worst16, kinds))
    print(msg)
    return msg


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 60.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
