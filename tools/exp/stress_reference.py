"""Randomised drop-in check against the UNMODIFIED reference running on the same GPU (models/blur_functions.py staged under
baseline/_ref by __graft_entry__.build(), or DIB_REFERENCE_ROOT): random lists of images (sizes, channel counts, fp32 / fp16),
random blurring flags, PSFs of every sweep cell, with and without the noise epilogue under the same numpy / torch seeds.
The exact-order path must reproduce the reference's tensors bit for bit, the default (tiled) path within 1e-5 (fp32) /
2e-2 (fp16: the reference rounds to half after every tap).
    python tools/exp/stress_reference.py [seconds] [seed]"""
import os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import detectinblur_b200.blur_functions as bf
import detectinblur_b200.psf_ops as ops
from detectinblur_b200.motion_blur import Trajectory


def run(budget=30.0, seed=0):
    import refshim
    ref_root = os.environ.get("DIB_REFERENCE_ROOT") or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "models")):
        msg = "skipped: no staged reference tree at %s" % ref_root
        print(msg)
        return msg
    refshim.install(ref_root)
    import models.blur_functions as rbf
    rng = random.Random(seed)
    np.random.seed(seed); random.seed(seed)
    dev = torch.device("cuda")
    cells = [(p, e) for p in (0.005, 0.001, 0.00005) for e in (1 / 25, 1 / 10, 1 / 5, 1 / 2, 1)]
    traj = np.stack([Trajectory(canvas=256, max_len=96, expl=p).fit().fit().x for p, _ in cells])
    pool = ops.rasterize_psfs(traj, np.array([e for _, e in cells]), dev, dtype=torch.float16)       # stored format: half
    t0 = time.time()
    n_cases, worst32, worst16 = 0, 0.0, 0.0
    while time.time() - t0 < budget:
        n = rng.randint(1, 4)
        half = rng.random() < 0.4
        noise = rng.random() < 0.3
        dt = torch.float16 if half else torch.float32
        imgs, dicts, psfs = [], [], []
        for _ in range(n):
            C, H, W = rng.choice([1, 3, 3]), rng.choice([65, 80, 131, 200]), rng.choice([65, 97, 224, 449, 460])
            imgs.append(torch.rand((C, H, W), device=dev).to(dt))
            blurring = rng.random() < 0.8
            dicts.append({"blurring": blurring})
            # engine.py:84: HalfTensor(psf) when blurring; a (1,)-shaped zero otherwise (transforms.py:456)
            psfs.append(pool[rng.randrange(len(cells))].to(dt) if blurring else torch.zeros(1, device=dev, dtype=dt))
        results = []
        for mode in ("reference", "exact", "default"):
            work = [t.clone() for t in imgs]
            np.random.seed(1000 + n_cases); torch.manual_seed(2000 + n_cases)
            if mode == "reference":
                rbf.blur_image_list(work, dicts, psfs, add_noise=noise, noise_level=0.01)
            else:
                bf.blur_image_list(work, dicts, psfs, add_noise=noise, noise_level=0.01, exact=(mode == "exact"))
            results.append(work)
        ref, exact, fast = results
        for k in range(n):
            r = ref[k].contiguous()
            if exact[k].shape != r.shape or not torch.equal(exact[k], r):
                print("EXACT MISMATCH", dict(case=n_cases, k=k, shape=tuple(imgs[k].shape), half=half, noise=noise, blurring=dicts[k]["blurring"],
                                             err=float((exact[k].float() - r.float()).abs().max()) if exact[k].shape == r.shape else -1))
                raise AssertionError("exact-order path differs from the reference loop")
            err = float((fast[k].float() - r.float()).abs().max())
            if half:
                worst16 = max(worst16, err)
            else:
                worst32 = max(worst32, err)
            if not err <= (2e-2 if half else 1e-5):
                print("TILED MISMATCH", dict(case=n_cases, k=k, shape=tuple(imgs[k].shape), half=half, noise=noise, err=err))
                raise AssertionError("default path differs from the reference loop")
        n_cases += 1
    msg = "ok: %d lists in %.0f s against the reference loop on CUDA, worst fp32 %.3g, worst fp16 %.3g" % (n_cases, time.time() - t0, worst32, worst16)
    print(msg)
    return msg


if __name__ == "__main__":
    run(float(sys.argv[1]) if len(sys.argv) > 1 else 30.0, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
