"""GPU tests of the host-side mirrors that sit either side of the blur kernel: BlurImage with the CUDA rasteriser,
deferred PSFs completed in the main process, the PSF class, and the fused blur -> normalize -> padded batch."""
import os
import random

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

from oracle import blur_oracle as bo  # noqa: E402
from oracle import psf_oracle as po  # noqa: E402
from tests.test_transforms_cpu import CONFIGS  # noqa: E402


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda")


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "transform_cases.npz"), allow_pickle=False)


def test_blur_image_on_the_fly_cuda_backend_matches_reference(cuda, golden_dir):
    from detectinblur_b200.transforms import BlurImage, complete_blur_dicts
    g = _golden(golden_dir)
    img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (96, 112, 3), dtype=np.uint8))
    checked = 0
    for n in range(int(g["n"])):
        ci, seed = (int(v) for v in g["cfg_%d" % n])
        kw = dict(CONFIGS[ci])
        if kw.get("use_stored_psfs") or not bool(g["blurring_%d" % n]):
            continue
        kw.setdefault("blur_image_in_transform", False)
        shape = tuple(int(v) for v in g["psf_shape_%d" % n])
        ref = np.zeros(int(np.prod(shape)))
        ref[g["psf_idx_%d" % n]] = g["psf_val_%d" % n]
        ref = ref.reshape(shape)
        for backend in ("cuda", "defer"):
            random.seed(seed)
            np.random.seed(seed)
            _, _, bd = BlurImage(psf_backend=backend, **kw)(img, None, {})
            if backend == "defer":
                assert bd["psf"] is None
                psfs = complete_blur_dicts([bd], cuda, dtype=torch.float16)
                assert psfs[0].dtype == torch.float16 and tuple(psfs[0].shape) == shape
                assert np.array_equal(psfs[0].cpu().numpy(), ref.astype(np.float16))
            assert bd["psf"].dtype == np.float64 and np.array_equal(bd["psf"], ref), (n, backend)
            theta, s1, s2 = g["scalars_%d" % n]
            assert bd["theta_rad"] == theta and bd["scale_factor_lambda1"] == s1 and bd["scale_factor_lambda2"] == s2
            assert random.random() == float(g["next_random_%d" % n])
        checked += 1
    assert checked >= 10


def test_psf_class_mirror(cuda, golden_dir):
    from detectinblur_b200.motion_blur import PSF, Trajectory
    g = np.load(os.path.join(golden_dir, "psf_cases.npz"), allow_pickle=False)
    for k in (0, 3, 8):
        expl, frac, seed = g["meta_%d" % k]
        np.random.seed(int(seed))
        tr = Trajectory(canvas=256, max_len=96, expl=expl).fit().fit()
        ps = PSF(canvas=256, trajectory=tr, fraction=[frac])
        out = ps.fit()
        raw = np.zeros(256 * 256)
        raw[g["raw_idx_%d" % k]] = g["raw_val_%d" % k]
        assert out is ps.PSFs and np.array_equal(ps.PSFs[0].ravel(), raw)
        ps.centerPSF()
        cen = np.zeros(256 * 256)
        cen[g["cen_idx_%d" % k]] = g["cen_val_%d" % k]
        assert np.array_equal(ps.PSFs[0].ravel(), cen)
        left, top, right, bottom = ps.findOffsets()
        ys, xs = np.nonzero(cen.reshape(256, 256))
        assert right == max(0, xs.max() - 127) and left == max(0, 127 - xs.min())
        assert bottom == max(0, ys.max() - 127) and top == max(0, 127 - ys.min())


FOURIER_CASES = ("noise", "smooth", "odd", "odd2", "gray", "small", "olddelta")


def test_blur_image_handler_matches_reference_fourier_path(cuda, golden_dir):
    """motion_blur.BlurImageHandler (the --cpu_blur blur) on the CUDA kernels against the reference's own outputs:
    even and odd sizes (the kernel origin moves to row 64), grey input, an image smaller than the kernel, oldDeltaPad.
    Tolerances: 1e-5 on result[0], <= 1 level on the truncated uint8 image."""
    from detectinblur_b200.motion_blur import BlurImageHandler
    g = np.load(os.path.join(golden_dir, "fourier_cases.npz"), allow_pickle=False)
    for name in FOURIER_CASES:
        h = BlurImageHandler(None, PSFs=[g["psf"].copy()], pillowImage=Image.fromarray(g["in_" + name]))
        assert h.blur_image(oldDeltaPad=(name == "olddelta")) is True
        res, u8 = h.result[0], np.array(h.pilImageResult)
        assert res.dtype == np.float32 and res.shape == g["res_" + name].shape, name
        # the Lanczos resize of the upscaled case amplifies differences slightly
        tol = 1e-5 if name != "small" else 2e-5
        assert np.abs(res - g["res_" + name]).max() <= tol, (name, np.abs(res - g["res_" + name]).max())
        assert u8.dtype == np.uint8 and u8.shape == g["u8_" + name].shape
        assert np.abs(u8.astype(int) - g["u8_" + name].astype(int)).max() <= 1, name
        assert (u8 != g["u8_" + name]).mean() < 5e-3, name


def test_baseline_config1_cpu_fourier_case_on_the_gpu(cuda):
    """BASELINE config 1, the reference's own CPU-runnable case: one 640x480 RGB uint8 image from
    default_rng(0), one PSF from the seeded generator (expl 0.005, low exposure), blurred through BlurImageHandler -- here
    on the CUDA kernels, against the CPU restatement of the Fourier path (pinned on the reference's outputs)."""
    from detectinblur_b200.motion_blur import BlurImageHandler
    from oracle import fourier_oracle as fo
    arr = np.random.default_rng(0).integers(0, 256, (480, 640, 3)).astype(np.uint8)
    np.random.seed(1337)
    random.seed(1337)
    fraction = random.choice([1 / 18, 1 / 10, 1 / 5])
    p16, _ = po.stored_psf(0.005, fraction, np.random)
    psf32 = po.crop128(p16).astype(np.float32)
    h = BlurImageHandler(None, PSFs=[psf32], pillowImage=Image.fromarray(arr))
    assert h.blur_image()
    want_f, want_u8 = fo.fourier_blur(arr, psf32)
    assert np.abs(h.result[0] - want_f).max() <= 1e-5
    got_u8 = np.array(h.pilImageResult)
    assert got_u8.shape == (480, 640, 3) and np.abs(got_u8.astype(int) - want_u8.astype(int)).max() <= 1
    assert (got_u8 != want_u8).mean() < 5e-3


def test_blur_image_handler_errors(cuda):
    from detectinblur_b200.motion_blur import BlurImageHandler
    with pytest.raises(Exception, match="Not correct path"):
        BlurImageHandler("/nonexistent.png", PSFs=[np.zeros((128, 128), np.float32)])
    with pytest.raises(AttributeError):
        BlurImageHandler(None, PSFs=None, pillowImage=Image.new("RGB", (200, 200)))


def test_cpu_blur_flag_blurs_on_the_gpu(cuda):
    """blur_image_in_transform=True (--cpu_blur): BlurImage returns the Fourier path's uint8 PIL image (edge padding,
    min-max stretch), produced by the CUDA kernel; checked against the CPU restatement of blur_image.py."""
    from detectinblur_b200.transforms import BlurImage
    from oracle import fourier_oracle as fo
    arr = np.random.default_rng(1).integers(0, 256, (150, 171, 3), dtype=np.uint8)
    img = Image.fromarray(arr)
    random.seed(11)
    np.random.seed(11)
    tr = BlurImage(prob=1.0, blur_type=0.005, blur_exposure=1 / 5, blur_image_in_transform=True, psf_backend="cuda")
    out, _, bd = tr(img, None, {})
    assert out.size == img.size and out.mode == "RGB" and tr.pilImageResult is out
    _, want = fo.fourier_blur(arr, bd["psf"].astype(np.float32))
    got = np.asarray(out)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    assert (got != want).mean() < 5e-3


def test_fused_blur_normalize_batch(cuda):
    from detectinblur_b200 import net_transforms as nt
    rng = np.random.default_rng(8)
    shapes = [(3, 97, 131), (3, 120, 100), (3, 70, 140)]
    imgs = [rng.random(s, dtype=np.float32) for s in shapes]
    np.random.seed(8)
    psfs = []
    for frac in (1 / 10, 1 / 5, 1 / 18):
        p16, _ = po.stored_psf(0.005, frac, np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    blur_dicts = [{"blurring": True}, {"blurring": False}, {"blurring": True}]
    means = np.array([[0.47, 0.45, 0.41], nt.CANONICAL_MEAN, [0.5, 0.4, 0.3]])
    stds = np.array([[0.21, 0.2, 0.22], nt.CANONICAL_STD, [0.25, 0.2, 0.3]])
    for exact in (True, False):
        il = nt.fused_blur_normalize([torch.from_numpy(a).to(cuda) for a in imgs], blur_dicts,
                                     [torch.from_numpy(p).to(cuda) for p in psfs], newMeans=means, newSTDs=stds, exact=exact)
        assert tuple(il.tensors.shape) == (3, 3, 128, 160) and il.image_sizes == [(97, 131), (120, 100), (70, 140)]
        for k, (a, bd) in enumerate(zip(imgs, blur_dicts)):
            base = bo.manual_blur(a, bo.normalize_psf(psfs[k])) if bd["blurring"] else a
            want = bo.normalize_image(base, means[k], stds[k])
            got = il.tensors[k, :, :a.shape[1], :a.shape[2]].cpu().numpy()
            if exact:
                assert np.array_equal(got, want), k
            else:
                assert np.abs(got - want).max() <= 1e-4, k
            assert il.tensors[k, :, a.shape[1]:, :].abs().max().item() == 0
            assert il.tensors[k, :, :, a.shape[2]:].abs().max().item() == 0
    # the reference-shaped module accepts the fused result unchanged
    tr = nt.GeneralizedRCNNTransform(800, 1333, nt.CANONICAL_MEAN, nt.CANONICAL_STD, normalize_images=False)
    out, _ = tr(il)
    assert out is il
    # and its own normalize + batch path agrees with the oracle's normalize on an unblurred list
    tr2 = nt.GeneralizedRCNNTransform(97, 131, nt.CANONICAL_MEAN, nt.CANONICAL_STD)
    one, _ = tr2([torch.from_numpy(imgs[0]).to(cuda)])
    assert np.array_equal(one.tensors[0, :, :97, :131].cpu().numpy(), bo.normalize_image(imgs[0], nt.CANONICAL_MEAN, nt.CANONICAL_STD))


def test_philox_noise_statistics(cuda):
    """In-kernel Philox noise (fast mode): clamp range, mean and variance of the injected noise."""
    import math
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops
    img = torch.full((3, 200, 300), 0.5, device=cuda)
    delta = torch.zeros(1, 128, 128, device=cuda)
    delta[0, 63, 63] = 1
    ts = ops.compact_taps(delta, normalize=False)
    sd = math.sqrt(0.004)
    for exact in (True, False):
        out = bf.blur_batch([img], ts, [0], noise_sd=[sd], philox_seed=1234, exact=exact)[0]
        assert out.min().item() >= 0 and out.max().item() <= 1
        d = (out - 0.5).double()
        assert abs(d.mean().item()) < 5e-4 and abs(d.std().item() - sd) < 1e-3
        again = bf.blur_batch([img], ts, [0], noise_sd=[sd], philox_seed=1234, exact=exact)[0]
        assert torch.equal(out, again)          # counter-based: same seed, same noise
        other = bf.blur_batch([img], ts, [0], noise_sd=[sd], philox_seed=99, exact=exact)[0]
        assert not torch.equal(out, other)


def test_estimator_mirror_and_bank_writer(cuda, golden_dir, tmp_path):
    import detectinblur_b200.engine_blur_estimator as est
    from detectinblur_b200 import psf_bank
    g = np.load(os.path.join(golden_dir, "estimator_cases.npz"), allow_pickle=False)
    psf = torch.from_numpy(g["psf"]).to(cuda)
    psfn = psf / psf.sum()
    for n in range(int(g["n"])):
        img = torch.from_numpy(g["img_%d" % n]).to(cuda)
        got = est.manual_blur(img, psfn, resize_images=True).cpu().numpy()
        want = g["out_%d" % n]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 3e-5, n       # GPU bilinear resize + tiled blur vs the reference on CPU
        lst = [img]
        est.blur_image_list(lst, [{"blurring": True}], [psf], resize_images=True)
        assert np.abs(lst[0].cpu().numpy() - want).max() <= 3e-5
    with pytest.raises(NotImplementedError):
        est.manual_blur(torch.rand(3, 900, 100, device=cuda), psfn, resize_images=True)
    # bank writer: byte-identical to what dataset_utils/generate_PSFs.py stores (oracle = the pinned restatement)
    dest = str(tmp_path) + "/"
    n_written = psf_bank.generate_psf_bank(dest, worker_index=1, num_workers=2, total_num_psfs=4, device=cuda)
    assert n_written == 3 * 5 * 2
    np.random.seed(1337 * 1)
    for p, expl in enumerate(psf_bank.PARAMS):
        for e, frac in enumerate(psf_bank.FRACTIONS):
            for index in (2, 3):
                want, _ = po.stored_psf(expl, frac, np.random)
                path = dest + "psfs/P%dE%d/I%06d" % (p + 1, e, index)
                got = np.load(open(path, "rb"))
                assert got.dtype == np.float16 and got.shape == (256, 256) and np.array_equal(got, want), path
                assert np.array_equal(psf_bank.load_stored_psf(dest + "psfs", p + 1, e, index), want[64:192, 64:192])


def test_resize_normalize_batch_matches_reference_transform(cuda, golden_dir):
    """dib_resize_batch (normalize + bilinear resize + zero-padded batch in one pass) against the reference's
    GeneralizedRCNNTransform.forward outputs; tolerance 5e-6 on normalised values (|x| <= 2.7; FMA contraction)."""
    from detectinblur_b200 import net_transforms as nt
    g = np.load(os.path.join(golden_dir, "resize_cases.npz"), allow_pickle=False)
    for n in range(int(g["n"])):
        mn, mx = (float(v) for v in g["minmax_%d" % n])
        imgs = [torch.from_numpy(g["img_%d_%d" % (n, k)]).to(cuda) for k in range(int(g["n_img_%d" % n]))]
        tr = nt.GeneralizedRCNNTransform(mn, mx, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], training=False)
        before = nt.launch_count()
        il, _ = tr(imgs, None, newMeans=g["means_%d" % n], newSTDs=g["stds_%d" % n])
        assert nt.launch_count() == before + 1                       # one kernel for the whole batch
        assert [list(s) for s in il.image_sizes] == g["sizes_%d" % n].tolist()
        got, want = il.tensors.cpu().numpy(), g["batch_%d" % n]
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 5e-6, (n, np.abs(got - want).max())
        # the padding is written by the kernel (the batch starts as torch.empty): exact zeros outside every image
        for k, (h, w) in enumerate(il.image_sizes):
            assert not got[k, :, h:, :].any() and not got[k, :, :, w:].any()
        # half I/O and no normalisation
        il16, _ = nt.GeneralizedRCNNTransform(mn, mx, None, None, training=False, normalize_images=False)([i.half() for i in imgs])
        ref = nt.GeneralizedRCNNTransform(mn, mx, None, None, training=False, normalize_images=False)([i.half().float().cpu() for i in imgs])[0]
        assert il16.tensors.dtype == torch.float16
        assert np.abs(il16.tensors.float().cpu().numpy() - ref.tensors.numpy()).max() <= 1e-3


def test_resize_pass_identity_scale_is_bit_exact_normalize(cuda):
    """An image already at the target scale leaves the fused pass exactly as `normalize` leaves it (IEEE subtract and
    divide, net_transforms.py:135-139): 3.2 M values against torch on the same device, several statistics."""
    from detectinblur_b200 import net_transforms as nt
    gen = torch.Generator().manual_seed(31)
    img = (torch.rand((3, 800, 1333), generator=gen) * 1.2 - 0.1).to(cuda)
    for mean, std in ((nt.CANONICAL_MEAN, nt.CANONICAL_STD), ([0.4695, 0.4461, 0.4068], [0.2005, 0.1962, 0.2006]),
                      ([0.0, 0.5, 1.0], [1.0, 0.003, 700.0])):
        tr = nt.GeneralizedRCNNTransform(800, 1333, mean, std, training=False)
        il, _ = tr([img])
        want = nt.normalize(img, mean, std)
        assert il.image_sizes == [(800, 1333)] and tuple(il.tensors.shape) == (1, 3, 800, 1344)
        assert torch.equal(il.tensors[0, :, :, :1333], want)
        assert not il.tensors[0, :, :, 1333:].any()


def test_fused_blur_normalize_with_resize(cuda):
    """blur at native size, then normalize + resize + batch: against oracle blur -> oracle transform."""
    from detectinblur_b200 import net_transforms as nt
    from oracle import resize_oracle as ro
    rng = np.random.default_rng(12)
    shapes = [(3, 97, 131), (3, 120, 100), (3, 70, 140)]
    imgs = [rng.random(s, dtype=np.float32) for s in shapes]
    np.random.seed(12)
    psfs = []
    for frac in (1 / 10, 1 / 5, 1 / 18):
        p16, _ = po.stored_psf(0.005, frac, np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    blur_dicts = [{"blurring": True}, {"blurring": False}, {"blurring": True}]
    means = np.array([[0.47, 0.45, 0.41], nt.CANONICAL_MEAN, [0.5, 0.4, 0.3]])
    stds = np.array([[0.21, 0.2, 0.22], nt.CANONICAL_STD, [0.25, 0.2, 0.3]])
    il = nt.fused_blur_normalize([torch.from_numpy(i).to(cuda) for i in imgs], blur_dicts,
                                 [torch.from_numpy(p).to(cuda) for p in psfs], newMeans=means, newSTDs=stds,
                                 min_size=160, max_size=200)
    blurred = [bo.manual_blur(imgs[k], bo.normalize_psf(psfs[k])) if blur_dicts[k]["blurring"] else imgs[k] for k in range(3)]
    want, sizes = ro.transform_forward(blurred, means, stds, 160, 200)
    assert [tuple(s) for s in il.image_sizes] == sizes
    got = il.tensors.cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5            # tiled blur (<= 1e-5 on [0,1]) scaled by 1/std ~ 5, plus the resize
    # at the requested scale already: the one-kernel path (blur epilogue writes the batch)
    same = nt.fused_blur_normalize([torch.from_numpy(imgs[0]).to(cuda)], [blur_dicts[0]], [torch.from_numpy(psfs[0]).to(cuda)],
                                   newMeans=means[:1], newSTDs=stds[:1], min_size=97, max_size=131)
    assert same.image_sizes == [(97, 131)]


def test_gpu_trajectories_match_restatement_and_reference_statistics(cuda):
    """dib_generate_trajectories: (1) equal to the CPU restatement of the same counter-based walk (<= 1e-8 px: libm vs
    CUDA sin / cos / log differ in the last bits); (2) the walk's invariants; (3) its statistics against the reference's
    MT19937 walk (host mirror, bit-identical to the reference): end-to-end displacement and bounding-box extent."""
    import detectinblur_b200.psf_ops as ops
    from detectinblur_b200.motion_blur.generate_trajectory import Trajectory
    expl = [0.005, 0.001, 0.00005, 0.005]
    x, big = ops.generate_trajectories(4, expl, 99, cuda, first_index=40, return_big_count=True)
    assert x.dtype == torch.complex128 and tuple(x.shape) == (4, 2000)
    xh = x.cpu().numpy()
    for k in range(4):
        want, wbig = po.trajectory_philox(99, 40 + k, expl[k])
        assert np.abs(xh[k] - want).max() <= 1e-8, k
        assert int(big[k]) == wbig
        assert xh[k][0] == complex(128, 128)
        assert abs(np.abs(np.diff(xh[k])).sum() - 96.0) < 1e-9           # every step has length max_len / (iters - 1)
    # a pure function of (seed, index): any batch split, explicit indices
    y = ops.generate_trajectories(2, expl[2:], 99, cuda, first_index=42)
    z = ops.generate_trajectories(2, [expl[3], expl[0]], 99, cuda, indices=[43, 40])
    assert torch.equal(y, x[2:]) and torch.equal(z[0], x[3]) and torch.equal(z[1], x[0])
    assert not torch.equal(ops.generate_trajectories(1, 0.005, 100, cuda, first_index=40)[0], x[0])
    # statistics against the reference walk
    for e, tol in ((0.005, 0.25), (0.00005, 0.1)):
        np.random.seed(4242)
        host = np.stack([Trajectory(canvas=256, max_len=96, expl=e).fit().x for _ in range(64)])
        gpu = ops.generate_trajectories(2048, e, 7, cuda).cpu().numpy()

        def stats(t):
            disp = np.abs(t[:, -1] - t[:, 0])
            ext = np.maximum(t.real.max(1) - t.real.min(1), t.imag.max(1) - t.imag.min(1))
            return disp, ext
        for hs, gs in zip(stats(host), stats(gpu)):
            se = np.sqrt(hs.var() / len(hs) + gs.var() / len(gs))
            assert abs(hs.mean() - gs.mean()) <= 4 * se + tol, (e, hs.mean(), gs.mean(), se)
    # device trajectories feed the rasteriser directly; a PSF's mass is its exposure fraction (generate_PSF.py:47-77)
    psf = ops.rasterize_psfs(x, [1 / 10, 1 / 5, 1 / 2, 1], cuda, canvas=256, center=True, out_side=256, dtype=torch.float64)
    sums = psf.sum(dim=(1, 2)).cpu().numpy()
    for k, f in enumerate((1 / 10, 1 / 5, 1 / 2, 1)):
        assert abs(sums[k] - po.time_weights(2000, f).sum() / 2000) < 1e-12


def test_blur_image_gpu_trajectory_backend(cuda):
    """BlurImage(trajectory_backend='cuda'): the walk is drawn on the GPU (numpy's stream untouched), directly or deferred."""
    from detectinblur_b200.transforms import BlurImage, complete_blur_dicts
    img = Image.fromarray(np.zeros((80, 90, 3), np.uint8))
    random.seed(3)
    np.random.seed(3)
    state = np.random.get_state()[1].copy()
    a = BlurImage(prob=1.0, blur_type=0.005, blur_exposure=1 / 5, blur_image_in_transform=False, psf_backend="cuda",
                  trajectory_backend="cuda", trajectory_seed=5)(img, None, {})[2]
    random.seed(3)
    b = BlurImage(prob=1.0, blur_type=0.005, blur_exposure=1 / 5, blur_image_in_transform=False, psf_backend="defer",
                  trajectory_backend="cuda", trajectory_seed=5)(img, None, {})[2]
    assert np.array_equal(np.random.get_state()[1], state)              # no numpy draws in this mode
    assert b["psf"] is None and b["deferred_psf"]["trajectory"] is None
    psfs = complete_blur_dicts([b], cuda)
    assert a["psf"].shape == (128, 128) and abs(a["psf"].sum() - 0.2) < 1e-3
    assert np.array_equal(a["psf"], b["psf"]) and a["theta_rad"] == b["theta_rad"]
    assert torch.equal(psfs[0].cpu(), torch.HalfTensor(b["psf"]))


def test_packed_bank_upload_and_writer(cuda, tmp_path):
    """Packed sparse bank on the device: dib_unpack_psfs expands a batch's taps into the dense PSFs the reference uploads
    one by one (engine.py:84); the packed writer stores the same PSFs as the dense writer; complete_blur_dicts uses it."""
    import detectinblur_b200.psf_ops as ops
    from detectinblur_b200 import psf_bank
    from detectinblur_b200.transforms import complete_blur_dicts
    dest = str(tmp_path) + "/"
    psf_bank.generate_psf_bank(dest, worker_index=1, num_workers=2, total_num_psfs=6, device=cuda, dense=True, packed=True)
    bank_dir = dest + "psfs"
    bank = psf_bank.PackedPsfBank(bank_dir)
    keys = [(p, e, i) for p in (1, 2, 3) for e in range(5) for i in (3, 4, 5)]
    dense_files = []
    for (p, e, i) in keys:
        full = np.load(open(bank_dir + "/P%dE%d/I%06d" % (p, e, i), "rb"))
        dense_files.append(full)
        assert np.array_equal(bank.dense(p, e, i).view(np.uint16), full[64:192, 64:192].view(np.uint16))
    want = np.stack([f[64:192, 64:192] for f in dense_files])
    for dt in (torch.float16, torch.float32, torch.float64):
        got = bank.upload(keys, cuda, dtype=dt)
        assert got.dtype == dt and tuple(got.shape) == (len(keys), 128, 128)
        assert torch.equal(got.cpu(), torch.from_numpy(want).to(dt))
    # whole canvas (crop_lo 0) straight through the C entry point's wrapper
    words = [np.asarray(bank.words(*k)) for k in keys[:4]]
    offs = np.concatenate([[0], np.cumsum([len(w) for w in words])])
    full = ops.unpack_psfs(np.concatenate(words), offs, cuda, crop_lo=0, out_side=256, dtype=torch.float16).cpu().numpy()
    assert np.array_equal(full.view(np.uint16), np.stack(dense_files[:4]).view(np.uint16))
    with pytest.raises(ValueError):
        ops.unpack_psfs(np.concatenate(words), offs[:-1], cuda)
    # the tap set built from the unpacked PSFs is the one built from the dense uploads
    a = ops.compact_taps(bank.upload(keys[:6], cuda, dtype=torch.float32), normalize=True)
    b = ops.compact_taps(torch.from_numpy(want[:6]).to(cuda).float(), normalize=True)
    for k in range(6):
        (ya, xa, wa), (yb, xb, wb) = a.taps(k), b.taps(k)
        assert np.array_equal(ya, yb) and np.array_equal(xa, xb) and np.array_equal(wa.view(np.uint32), wb.view(np.uint32))
    # complete_blur_dicts sends stored PSFs up as taps (blur_dicts as BlurImage leaves them, tests/test_transforms_cpu.py)
    dicts = [{"blurring": True, "psf": bank.dense(*k), "stored_psf_source": (bank_dir,) + k} for k in keys[:3]]
    dicts.insert(1, {"blurring": False, "psf": [0]})
    dicts.append({"blurring": True, "psf": want[7]})                        # no bank source: dense upload
    psfs = complete_blur_dicts(dicts, cuda)
    for bd, t in zip(dicts, psfs):
        if bd["blurring"]:
            assert t.dtype == torch.float16 and torch.equal(t.cpu(), torch.HalfTensor(bd["psf"]))
        else:
            assert tuple(t.shape) == (1,)


def test_eval_sweep_shapes_tiled_vs_exact(cuda):
    """BASELINE config 4: P in {0.005, 0.001, 0.00005} x E in {1/25, 1/10, 1/5, 1/2, 1} (evaluate.py:299-300) on
    COCO-val-like shapes: rasterise on the GPU, blur with the tiled kernel, compare with the exact-order kernel."""
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops
    from detectinblur_b200.motion_blur import Trajectory
    shapes = [(480, 640), (427, 640), (640, 480), (640, 427), (375, 500), (500, 375), (333, 500)]
    np.random.seed(2024)
    traj, frac = [], []
    for expl in (0.005, 0.001, 0.00005):
        for f in (1 / 25, 1 / 10, 1 / 5, 1 / 2, 1):
            traj.append(Trajectory(canvas=256, max_len=96, expl=expl).fit().fit().x)
            frac.append(f)
    psfs = ops.rasterize_psfs(np.stack(traj), frac, cuda, dtype=torch.float32)
    ts = ops.compact_taps(psfs, normalize=True)
    g = torch.Generator().manual_seed(7)
    imgs = [torch.rand((3,) + shapes[k % len(shapes)], generator=g).to(cuda) for k in range(15)]
    fast = bf.blur_batch(imgs, ts, list(range(15)), exact=False)
    exact = bf.blur_batch(imgs, ts, list(range(15)), exact=True)
    for k in range(15):
        assert (fast[k].double() - exact[k].double()).abs().max().item() <= 1e-5, (k, ts.counts[k])
