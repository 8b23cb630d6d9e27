"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only (the reference tree does not travel to the GPU box):

    python tools/make_golden.py

Every fixture stores the inputs (or the seeds that regenerate them) and the reference's outputs.  The
reference has no tests of its own, so these vectors are what pins the oracle (oracle/__init__.py).
"""
import os
import random
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()

import torch  # noqa: E402
from PIL import Image  # noqa: E402
from motion_blur.generate_trajectory import Trajectory  # noqa: E402
from motion_blur.generate_PSF import PSF  # noqa: E402
from motion_blur.blur_image import BlurImageHandler  # noqa: E402
import models.blur_functions as ref_bf  # noqa: E402
import transforms as ref_T  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
PARAMS = [0.005, 0.001, 0.00005]
FRACTIONS = [1 / 18, 1 / 10, 1 / 5, 1 / 2, 1]


def ref_psf(expl, fraction, seed, center=True):
    np.random.seed(seed)
    random.seed(seed)
    tr = Trajectory(canvas=256, max_len=96, expl=expl).fit().fit()
    ps = PSF(canvas=256, trajectory=tr, fraction=[fraction])
    ps.fit()
    raw = ps.PSFs[0].copy()
    if center:
        ps.centerPSF()
    return tr.x.copy(), raw, ps.PSFs[0].copy()


def sparse(a):
    idx = np.flatnonzero(a)
    return idx.astype(np.int32), a.ravel()[idx]


def gen_psf_cases():
    cases = {}
    k = 0
    for pi, p in enumerate(PARAMS):
        for fi in (0, 2, 4) if pi != 1 else (1, 3):
            x, raw, cen = ref_psf(p, FRACTIONS[fi], 1000 + k)
            ri, rv = sparse(raw)
            ci, cv = sparse(cen)
            cases["x_%d" % k] = x
            cases["meta_%d" % k] = np.array([p, FRACTIONS[fi], 1000 + k], dtype=np.float64)
            cases["raw_idx_%d" % k], cases["raw_val_%d" % k] = ri, rv
            cases["cen_idx_%d" % k], cases["cen_val_%d" % k] = ci, cv
            k += 1
    # the evaluation sweep's 1/25 exposure (evaluate.py:300)
    x, raw, cen = ref_psf(0.005, 1 / 25, 1000 + k)
    ri, rv = sparse(raw)
    ci, cv = sparse(cen)
    cases["x_%d" % k] = x
    cases["meta_%d" % k] = np.array([0.005, 1 / 25, 1000 + k], dtype=np.float64)
    cases["raw_idx_%d" % k], cases["raw_val_%d" % k] = ri, rv
    cases["cen_idx_%d" % k], cases["cen_val_%d" % k] = ci, cv
    k += 1
    cases["n"] = np.array(k)
    np.savez_compressed(os.path.join(OUT, "psf_cases.npz"), **cases)
    print("psf cases:", k)


def gen_blur_cases():
    """manual_blur / blur_image_list on CPU tensors, fp32 and fp16."""
    cases = {}
    specs = []
    # (C, H, W, psf kind, seed)
    specs.append((3, 70, 90, ("gen", 0.005, 1 / 10), 11))
    specs.append((3, 65, 65, ("gen", 0.001, 1 / 5), 12))
    specs.append((1, 80, 72, ("gen", 0.005, 1 / 18), 13))       # C == 1 -> 2-D result
    specs.append((3, 40, 50, ("gen", 0.005, 1 / 5), 14))        # zero-pad mode (both sides < 64)
    specs.append((3, 63, 100, ("gen", 0.00005, 1 / 2), 15))     # zero-pad mode (H < 64 only)
    specs.append((3, 97, 131, ("gen", 0.00005, 1), 16))         # long exposure
    specs.append((3, 80, 77, ("edge",), 17))                    # taps on PSF rows/cols 0 and 127 (roll wrap quirk)
    specs.append((3, 140, 150, ("gen256", 0.001, 1 / 2), 18))   # 256 branch (--dont_center_psf), replicate pad
    specs.append((2, 30, 20, ("gen256", 0.005, 1 / 5), 19))     # 256 branch on a tiny image
    specs.append((3, 66, 300, ("gen", 0.001, 1), 20))
    for n, (C, H, W, kind, seed) in enumerate(specs):
        rng = np.random.default_rng(seed)
        img = rng.random((C, H, W), dtype=np.float32)
        if kind[0] == "gen":
            _, _, cen = ref_psf(kind[1], kind[2], seed)
            psf = cen.astype(np.float16)[64:192, 64:192].astype(np.float32)
        elif kind[0] == "gen256":
            _, raw, _ = ref_psf(kind[1], kind[2], seed, center=False)
            psf = raw.astype(np.float16).astype(np.float32)
        else:
            psf = np.zeros((128, 128), np.float32)
            for (y, x) in ((0, 0), (0, 127), (127, 0), (127, 127), (63, 63), (127, 64), (64, 127), (0, 60), (5, 0)):
                psf[y, x] = rng.random()
        for dt, tdt in (("f32", torch.float32), ("f16", torch.float16)):
            t_img = torch.from_numpy(img).to(tdt)
            t_psf = torch.from_numpy(psf).to(tdt)
            t_psf_n = t_psf / t_psf.sum()
            out = ref_bf.manual_blur(t_img, t_psf_n)
            cases["out_%s_%d" % (dt, n)] = out.contiguous().numpy()
            cases["psfn_%s_%d" % (dt, n)] = t_psf_n.numpy()
        cases["img_%d" % n] = img
        cases["psf_%d" % n] = psf
    cases["n"] = np.array(len(specs))

    # noise epilogue (blur_functions.py:72-74) with the draws captured
    rng = np.random.default_rng(99)
    img = rng.random((3, 70, 75), dtype=np.float32)
    _, _, cen = ref_psf(0.005, 1 / 5, 99)
    psf = cen.astype(np.float16)[64:192, 64:192].astype(np.float32)
    psf_n = torch.from_numpy(psf) / torch.from_numpy(psf).sum()
    np.random.seed(5)
    torch.manual_seed(5)
    out = ref_bf.manual_blur(torch.from_numpy(img), psf_n, add_noise=True, noise_level=0.01)
    np.random.seed(5)
    torch.manual_seed(5)
    noise_var = np.random.uniform(0.00000001, 0.01)
    noise = torch.randn(3, 70, 75)
    cases["noise_img"], cases["noise_psf"], cases["noise_out"] = img, psf, out.contiguous().numpy()
    cases["noise_var"], cases["noise_draw"] = np.array(noise_var), noise.numpy()

    # blur_image_list with a non-blurring entry (identity preserved) and its own normalisation
    imgs = [torch.from_numpy(rng.random((3, 66, 68), dtype=np.float32)) for _ in range(3)]
    keep = imgs[1]
    psfs = []
    for s in (31, 32, 33):
        _, _, cen = ref_psf(0.005, 1 / 10, s)
        psfs.append(torch.from_numpy(cen.astype(np.float16)[64:192, 64:192].astype(np.float32)))
    psfs[1] = torch.tensor([0.0])
    bds = [{"blurring": True}, {"blurring": False}, {"blurring": True}]
    cases["list_in"] = np.stack([i.numpy().copy() for i in imgs])
    cases["list_psf0"], cases["list_psf2"] = psfs[0].numpy(), psfs[2].numpy()
    ref_bf.blur_image_list(imgs, bds, psfs)
    assert imgs[1] is keep
    cases["list_out"] = np.stack([i.contiguous().numpy() for i in imgs])
    np.savez_compressed(os.path.join(OUT, "blur_cases.npz"), **cases)
    print("blur cases:", len(specs))

    # H == 64 must raise in the reference (reflect pad 64 >= dim)
    try:
        ref_bf.manual_blur(torch.rand(3, 64, 80), psf_n)
        raise SystemExit("expected the reference to raise on H == 64")
    except RuntimeError:
        pass


def gen_transform_cases():
    """BlurImage.__call__ in gpu mode (blur_image_in_transform=False): RNG consumption + blur_dict contents."""
    bank = tempfile.mkdtemp(prefix="dib_bank_")
    img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (96, 112, 3), dtype=np.uint8))
    configs = [
        dict(prob=0.9, use_stored_psfs=False, low_exposure=False, high_exposure=False),
        dict(prob=0.75, blur_type=0.005, use_stored_psfs=False, low_exposure=True),
        dict(prob=1.0, blur_type=0.00005, use_stored_psfs=False, high_exposure=True),
        dict(prob=1.0, blur_type=0.001, blur_exposure=1 / 25, use_stored_psfs=False),
        dict(prob=0.75, blur_type=1, use_stored_psfs=True, low_exposure=True),
        dict(prob=1.0, blur_type=3, use_stored_psfs=True, high_exposure=True),
        dict(prob=0.9, use_stored_psfs=True),
        dict(prob=1.0, use_stored_psfs=False, dont_center_psf=True, blur_type=0.001, low_exposure=True),
        dict(prob=0.0),
    ]
    cases = {}
    bank_files = {}
    n = 0
    for ci, cfg in enumerate(configs):
        for seed in (7, 8, 9):
            kw = dict(cfg)
            kw.setdefault("blur_image_in_transform", False)
            if kw.get("use_stored_psfs"):
                kw["stored_psf_directory"] = bank
            for attempt in range(4):
                random.seed(seed)
                np.random.seed(seed)
                tf = ref_T.BlurImage(**kw)
                try:
                    _, _, bd = tf(img, None, {})
                    break
                except FileNotFoundError as e:
                    # materialise exactly the bank file the reference asked for
                    path = e.filename
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    cell = os.path.basename(os.path.dirname(path))
                    p, f = int(cell[1]), int(cell[3])
                    st = np.random.get_state()
                    _, _, cen = ref_psf(PARAMS[p - 1], FRACTIONS[f], 5000 + len(bank_files))
                    np.random.set_state(st)
                    with open(path, "wb") as fh:
                        np.save(fh, cen.astype(np.float16))
                    bank_files[os.path.relpath(path, bank)] = cen.astype(np.float16)
            else:
                raise SystemExit("bank materialisation failed")
            cases["cfg_%d" % n] = np.array([ci, seed])
            cases["blurring_%d" % n] = np.array(bool(bd["blurring"]))
            psf = np.asarray(bd["psf"])
            cases["psf_shape_%d" % n] = np.array(psf.shape)
            cases["psf_dtype_%d" % n] = np.array(str(psf.dtype))
            if bd["blurring"]:
                idx, val = sparse(psf)
                cases["psf_idx_%d" % n], cases["psf_val_%d" % n] = idx, val
            cases["scalars_%d" % n] = np.array([bd["theta_rad"], bd["scale_factor_lambda1"], bd["scale_factor_lambda2"]],
                                               dtype=np.float64)
            cases["indices_%d" % n] = np.array([-99 if bd["param_index"] is None else bd["param_index"],
                                                -99 if bd["fraction_index"] is None else bd["fraction_index"]])
            # state of python's RNG after the call pins how many draws were consumed
            cases["next_random_%d" % n] = np.array(random.random())
            n += 1
    cases["n"] = np.array(n)
    names = sorted(bank_files)
    cases["bank_names"] = np.array(names)
    for i, nm in enumerate(names):
        idx, val = sparse(bank_files[nm])
        cases["bank_idx_%d" % i], cases["bank_val_%d" % i] = idx, val
    shutil.rmtree(bank)
    np.savez_compressed(os.path.join(OUT, "transform_cases.npz"), **cases)
    print("transform cases:", n, "bank files:", len(names))
    return configs


def gen_fourier_case():
    rng = np.random.default_rng(0)
    arr = rng.integers(0, 256, (160, 200, 3), dtype=np.uint8)
    # a smooth image as well: the min-max stretch depends on image contrast
    yy, xx = np.mgrid[0:160, 0:200]
    smooth = np.stack([40 + 150 * xx / 199, 60 + 100 * yy / 159, 90 + 60 * np.sin(xx / 17.0)], axis=2).astype(np.uint8)
    # odd sizes move the zero-padded kernel's centre (blur_image.py:119-123); a 2-D image is replicated to RGB (:91-97);
    # an image smaller than the kernel is upscaled first and scaled back afterwards (:56-69, :142-143)
    odd = rng.integers(0, 256, (161, 203, 3), dtype=np.uint8)
    odd2 = rng.integers(0, 256, (150, 131, 3), dtype=np.uint8)
    gray = rng.integers(0, 256, (140, 150), dtype=np.uint8)
    small = rng.integers(0, 256, (90, 100, 3), dtype=np.uint8)
    _, _, cen = ref_psf(0.005, 1 / 10, 1337)
    psf = cen[64:192, 64:192].astype(np.float32)
    cases = {"psf": psf}
    for name, a, kw in (("noise", arr, {}), ("smooth", smooth, {}), ("odd", odd, {}), ("odd2", odd2, {}), ("gray", gray, {}),
                        ("small", small, {}), ("olddelta", rng.integers(0, 256, (171, 171, 3), dtype=np.uint8), {"oldDeltaPad": True})):
        h = BlurImageHandler(None, PSFs=[psf.copy()], pillowImage=Image.fromarray(a))
        assert h.blur_image(**kw)
        cases["in_" + name] = a
        cases["res_" + name] = h.result[0]
        cases["u8_" + name] = np.array(h.pilImageResult)
    np.savez_compressed(os.path.join(OUT, "fourier_cases.npz"), **cases)
    print("fourier cases: 7")


def gen_normalize_case():
    from models.net_transforms import GeneralizedRCNNTransform
    tr = GeneralizedRCNNTransform(800, 1333, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    rng = np.random.default_rng(4)
    img = rng.random((3, 37, 41), dtype=np.float32)
    mean = np.array([0.4695, 0.4461, 0.4068])
    std = np.array([0.2087, 0.2043, 0.2088]) * 0.229 / 0.2384
    out = tr.normalize(torch.from_numpy(img), mean, std).numpy()
    np.savez_compressed(os.path.join(OUT, "normalize_case.npz"), img=img, mean=mean, std=std, out=out)
    print("normalize case: 1")


def gen_resize_cases():
    """GeneralizedRCNNTransform.forward in eval mode (normalize with per-image statistics, bilinear resize, zero-padded
    batch; net_transforms.py:82-133) on small COCO-shaped images: up- and down-scaling, landscape and portrait."""
    from models.net_transforms import GeneralizedRCNNTransform
    rng = np.random.default_rng(21)
    cases = {}
    specs = [((100, 140), [(3, 60, 80), (3, 80, 60), (3, 47, 83)]),       # upscale, max_size binds for the third
             ((64, 96), [(3, 120, 161), (3, 97, 75)]),                      # downscale
             ((90, 120), [(3, 90, 120), (3, 90, 101)])]                     # identity scale and a near-identity one
    for n, ((mn, mx), shapes) in enumerate(specs):
        tr = GeneralizedRCNNTransform(mn, mx, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], training=False)
        imgs = [rng.random(s, dtype=np.float32) for s in shapes]
        means = np.stack([np.array([0.485, 0.456, 0.406]) + 0.02 * rng.standard_normal(3) for _ in shapes])
        stds = np.stack([np.array([0.229, 0.224, 0.225]) * (1 + 0.1 * rng.random(3)) for _ in shapes])
        il, _ = tr([torch.from_numpy(i) for i in imgs], None, newMeans=means, newSTDs=stds)
        cases["minmax_%d" % n] = np.array([mn, mx])
        cases["n_img_%d" % n] = len(shapes)
        for k, im in enumerate(imgs):
            cases["img_%d_%d" % (n, k)] = im
        cases["means_%d" % n] = means
        cases["stds_%d" % n] = stds
        cases["batch_%d" % n] = il.tensors.numpy()
        cases["sizes_%d" % n] = np.array(il.image_sizes)
    cases["n"] = len(specs)
    np.savez_compressed(os.path.join(OUT, "resize_cases.npz"), **cases)
    print("resize cases: %d" % len(specs))


def gen_estimator_cases():
    """engine_blur_estimator.manual_blur with resize_images (:27-70).  The module itself cannot be imported here (it pulls in
    pycocotools), so the two function definitions are executed from the reference file at run time; nothing is copied."""
    import math
    src = open(os.path.join(refshim.install(), "engine_blur_estimator.py")).read().split("\n")
    ns = {"torch": torch, "math": math}
    exec("\n".join(src[26:79]), ns)
    cases = {}
    _, _, cen = ref_psf(0.005, 1 / 10, 77)
    psf = cen.astype(np.float16)[64:192, 64:192].astype(np.float32)
    psfn = torch.from_numpy(psf) / torch.from_numpy(psf).sum()
    rng = np.random.default_rng(77)
    for n, (C, H, W) in enumerate([(3, 120, 200), (3, 200, 120), (1, 90, 90)]):
        img = rng.random((C, H, W), dtype=np.float32)
        out = ns["manual_blur"](torch.from_numpy(img), psfn, True)
        cases["img_%d" % n], cases["out_%d" % n] = img, out.contiguous().numpy()
    cases["psf"] = psf
    cases["n"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "estimator_cases.npz"), **cases)
    print("estimator cases: 3")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gens = [gen_psf_cases, gen_blur_cases, gen_transform_cases, gen_fourier_case, gen_normalize_case, gen_resize_cases, gen_estimator_cases]
    only = sys.argv[1:]          # e.g. `python tools/make_golden.py gen_fourier_case` regenerates one fixture
    for g in gens:
        if not only or g.__name__ in only:
            g()
