"""GPU parity tests: the CUDA path (through the C ABI, via detectinblur_b200) against the golden fixtures made by
the unmodified reference and against the CPU oracle on seeded inputs.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import blur_oracle as bo  # noqa: E402
from oracle import psf_oracle as po  # noqa: E402

TOL_FP32 = 1e-5   # north_star: blurred pixels within max-abs 1e-5 of the reference's fp32 GPU loop


@pytest.fixture(scope="module")
def dib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import detectinblur_b200.blur_functions as bf
    import detectinblur_b200.psf_ops as ops
    return bf, ops


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def test_taps_bit_exact_vs_golden_and_oracle(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    for n in range(int(g["n"])):
        for dt, tdt in (("f32", torch.float32), ("f16", torch.float16)):
            psfn = g["psfn_%s_%d" % (dt, n)]
            ts = ops.compact_taps(_cuda(psfn), normalize=False)
            ys, xs, ws = ts.taps(0)
            oy, ox, ow = bo.compact_taps(psfn)
            assert np.array_equal(ys, oy) and np.array_equal(xs, ox), (n, dt)
            assert np.array_equal(ws, ow.astype(np.float32)), (n, dt)
            assert ts.counts[0] == len(oy)


def test_normalisation_matches_torch_cuda_bit_exact(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    for n in range(int(g["n"])):
        for tdt in (torch.float32, torch.float16):
            psf = _cuda(g["psf_%d" % n], tdt)
            ref = (psf / psf.sum())                       # blur_functions.py:98 on the device
            ts = ops.compact_taps(psf, normalize=True)
            ys, xs, ws = ts.taps(0)
            nz = ref.nonzero(as_tuple=False).cpu().numpy()
            assert np.array_equal(nz[:, 0], ys) and np.array_equal(nz[:, 1], xs), (n, tdt)
            assert np.array_equal(ref[ref != 0].float().cpu().numpy(), ws), (n, tdt)
            # and the oracle's restatement
            on = bo.normalize_psf(g["psf_%d" % n].astype(np.float16 if tdt == torch.float16 else np.float32))
            assert np.array_equal(bo.compact_taps(on)[2].astype(np.float32), ws), (n, tdt)


def test_blur_golden_exact_and_tiled(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    for n in range(int(g["n"])):
        img = g["img_%d" % n]
        for dt, tdt in (("f32", torch.float32), ("f16", torch.float16)):
            psfn = _cuda(g["psfn_%s_%d" % (dt, n)])
            ref = g["out_%s_%d" % (dt, n)]
            got = bf.manual_blur(_cuda(img, tdt), psfn, exact=True)
            assert tuple(got.shape) == ref.shape, (n, dt)
            assert np.array_equal(got.cpu().numpy(), ref), "exact kernel, case %d %s" % (n, dt)
            if tdt == torch.float32:
                fast = bf.manual_blur(_cuda(img, tdt), psfn, exact=False).cpu().numpy()
                err = np.abs(fast.astype(np.float64) - ref.astype(np.float64)).max()
                assert err <= TOL_FP32, "tiled kernel, case %d: max abs err %g" % (n, err)


def test_noise_epilogue_golden(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    psfn = bo.normalize_psf(g["noise_psf"])
    ts = ops.compact_taps(_cuda(psfn), normalize=False)
    import math
    for exact in (True, False):
        out = bf.blur_batch([_cuda(g["noise_img"])], ts, [0], noise=[_cuda(g["noise_draw"])],
                            noise_sd=[math.sqrt(float(g["noise_var"]))], exact=exact)[0].cpu().numpy()
        if exact:
            assert np.array_equal(out, g["noise_out"])
        else:
            assert np.abs(out - g["noise_out"]).max() <= TOL_FP32
        assert out.min() >= 0 and out.max() <= 1


def test_blur_image_list_golden(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    for exact in (True, False):
        imgs = [_cuda(a) for a in g["list_in"]]
        keep = imgs[1]
        psfs = [_cuda(g["list_psf0"]), torch.tensor([0.0]).cuda(), _cuda(g["list_psf2"])]
        r = bf.blur_image_list(imgs, [{"blurring": True}, {"blurring": False}, {"blurring": True}], psfs, exact=exact)
        assert r is None
        assert imgs[1] is keep
        got = np.stack([i.cpu().numpy() for i in imgs])
        if exact:
            assert np.array_equal(got, g["list_out"])
        else:
            assert np.abs(got - g["list_out"]).max() <= TOL_FP32


def test_reflect_needs_side_above_64(dib):
    bf, ops = dib
    psf = torch.zeros(128, 128).cuda()
    psf[63, 63] = 1
    with pytest.raises(RuntimeError):
        bf.manual_blur(torch.rand(3, 64, 80).cuda(), psf)
    with pytest.raises(RuntimeError):
        bf.manual_blur(torch.rand(3, 80, 64).cuda(), psf)
    bf.manual_blur(torch.rand(3, 65, 65).cuda(), psf)
    with pytest.raises(RuntimeError):
        bf.manual_blur(torch.rand(3, 80, 80), psf)    # CPU tensors are refused: there is no CPU path


@pytest.mark.parametrize("shape", [(3, 200, 300), (3, 81, 225), (1, 65, 449), (4, 161, 230), (3, 480, 640)])
def test_tiled_vs_oracle_seeded(dib, shape):
    """Sizes the oracle finishes in seconds; odd shapes straddle tile (36 x 448) and warp-block (6 x 224) boundaries."""
    bf, ops = dib
    rng = np.random.default_rng(sum(shape))
    img = rng.random(shape, dtype=np.float32)
    np.random.seed(shape[1])
    psf16, _ = po.stored_psf(0.001, 1 / 2, np.random)
    psf = po.crop128(psf16).astype(np.float32)
    psfn = bo.normalize_psf(psf)
    want = bo.manual_blur(img, psfn)
    got_exact = bf.manual_blur(_cuda(img), _cuda(psfn), exact=True).cpu().numpy()
    assert np.array_equal(got_exact, want)
    got = bf.manual_blur(_cuda(img), _cuda(psfn), exact=False).cpu().numpy()
    assert np.abs(got.astype(np.float64) - want).max() <= TOL_FP32
    # both row stores of the tiled kernel: the default result has 16-byte-aligned rows (a view of a wider allocation
    # when W % 4 != 0), a caller-provided contiguous destination has not -- same values either way
    ts = ops.compact_taps(_cuda(psfn), normalize=False)
    dflt = bf.blur_batch([_cuda(img)], ts, [0])[0]
    assert dflt.data_ptr() % 16 == 0 and dflt.stride(1) % 4 == 0 and tuple(dflt.shape) == shape
    packed = torch.empty(shape, device="cuda")
    res = bf.blur_batch([_cuda(img)], ts, [0], outs=[packed])[0]
    assert res is packed and packed.is_contiguous()
    odd = torch.empty((shape[0], shape[1], shape[2] + 1), device="cuda")[:, :, 1:]      # rows start 4 bytes off alignment
    bf.blur_batch([_cuda(img)], ts, [0], outs=[odd])
    assert torch.equal(dflt, packed) and torch.equal(dflt, odd)
    assert np.array_equal(dflt.cpu().numpy().reshape(want.shape), got)


@pytest.mark.parametrize("shape", [(3, 200, 300), (1, 65, 449), (2, 130, 1000), (3, 289, 331)])
def test_zero_padding_mode_both_kernels(dib, shape):
    """Zero padding at any size (what the Fourier-path mirror asks for): exact-order kernel bit-exact, tiled within 1e-5."""
    bf, ops = dib
    from detectinblur_b200 import _lib
    rng = np.random.default_rng(sum(shape) + 1)
    img = rng.random(shape, dtype=np.float32)
    np.random.seed(shape[2])
    psf16, _ = po.stored_psf(0.001, 1 / 2, np.random)
    psfn = bo.normalize_psf(po.crop128(psf16).astype(np.float32))
    want = bo.manual_blur(img, psfn, pad_mode=bo.PAD_ZERO).reshape(shape)
    ts = ops.compact_taps(_cuda(psfn), normalize=False)
    got_exact = bf.blur_batch([_cuda(img)], ts, [0], pad_mode=_lib.PAD_ZERO128, exact=True)[0].cpu().numpy()
    assert np.array_equal(got_exact, want)
    got = bf.blur_batch([_cuda(img)], ts, [0], pad_mode=_lib.PAD_ZERO128, exact=False)[0].cpu().numpy()
    assert np.abs(got.astype(np.float64) - want).max() <= TOL_FP32
    assert bf.launch_count() > 0


def test_full_size_properties(dib):
    """BASELINE config sizes (3 x 800 x 1333): size-independent properties + tiled vs exact-order kernel on device."""
    bf, ops = dib
    g = torch.Generator(device="cpu").manual_seed(1337)
    batch = torch.rand((4, 3, 800, 1333), generator=g).cuda()
    np.random.seed(0)
    psfs = []
    for frac, expl in ((1 / 18, 0.005), (1 / 5, 0.005), (1 / 2, 0.00005), (1, 0.00005)):
        p16, _ = po.stored_psf(expl, frac, np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    psfs_t = _cuda(np.stack(psfs))
    ts = ops.compact_taps(psfs_t, normalize=True)
    imgs = [batch[k] for k in range(4)]
    fast = bf.blur_batch(imgs, ts, [0, 1, 2, 3], exact=False)
    exact = bf.blur_batch(imgs, ts, [0, 1, 2, 3], exact=True)
    for k in range(4):
        err = (fast[k].double() - exact[k].double()).abs().max().item()
        assert err <= TOL_FP32, (k, err, ts.counts[k])
    # identity PSF: a single tap at the centre returns the image bit for bit
    delta = torch.zeros(1, 128, 128).cuda()
    delta[0, 63, 63] = 1
    tsd = ops.compact_taps(delta, normalize=False)
    assert torch.equal(bf.blur_batch([imgs[0]], tsd, [0])[0], imgs[0])
    # shifted delta: out[i, j] = img[i - dy, j - dx] in the interior (convolution orientation of the reference)
    sh = torch.zeros(1, 128, 128).cuda()
    sh[0, 63 + 5, 63 - 7] = 1
    tss = ops.compact_taps(sh, normalize=False)
    o = bf.blur_batch([imgs[1]], tss, [0])[0]
    assert torch.equal(o[:, 70:700, 70:1200], imgs[1][:, 65:695, 77:1207])
    # linearity: blur(a + b) = blur(a) + blur(b) within rounding; constant image stays constant (weights sum to 1)
    a, b = imgs[2] * 0.5, imgs[3] * 0.5
    lhs = bf.blur_batch([a + b], ts, [3])[0]
    rhs = bf.blur_batch([a], ts, [3])[0] + bf.blur_batch([b], ts, [3])[0]
    assert (lhs - rhs).abs().max().item() <= 2e-6
    const = torch.full((3, 800, 1333), 0.625).cuda()
    oc = bf.blur_batch([const], ts, [2])[0]
    assert (oc - 0.625).abs().max().item() <= 2e-6


def test_fused_normalize_and_pitched_output(dib):
    bf, ops = dib
    rng = np.random.default_rng(5)
    img = rng.random((3, 97, 131), dtype=np.float32)
    np.random.seed(5)
    p16, _ = po.stored_psf(0.005, 1 / 5, np.random)
    psfn = bo.normalize_psf(po.crop128(p16).astype(np.float32))
    mean = [0.485, 0.456, 0.406]
    std = [0.229, 0.224, 0.225]
    want = bo.normalize_image(bo.manual_blur(img, psfn), mean, std)
    ts = ops.compact_taps(_cuda(psfn), normalize=False)
    batch = torch.zeros((1, 3, 128, 160)).cuda()          # zero-padded batch as batch_images builds (net_transforms.py:218-249)
    out_view = batch[0, :, :97, :131]
    for exact in (True, False):
        batch.zero_()
        bf.blur_batch([_cuda(img)], ts, [0], outs=[out_view], mean=[mean], std=[std], exact=exact)
        got = batch[0, :, :97, :131].cpu().numpy()
        if exact:
            assert np.array_equal(got, want)
        else:
            assert np.abs(got - want).max() <= 1e-4     # 1e-5 on the blur, divided by std ~ 0.22
        assert batch[0, :, 97:, :].abs().max().item() == 0 and batch[0, :, :, 131:].abs().max().item() == 0


def test_psf_metadata(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    for n in range(int(g["n"])):
        psf = g["psf_%d" % n]
        if psf.shape[0] != 128:
            continue
        psfn = bo.normalize_psf(psf)
        ts = ops.compact_taps(_cuda(psf), normalize=True)
        assert ts.tap_extents(0) == bo.tap_extents(psfn)
        th, s1, s2 = ts.psf_pca(0)
        oth, os1, os2 = bo.psf_pca(psf)
        assert abs(th - oth) < 1e-9 and abs(s1 - os1) < 1e-12 and abs(s2 - os2) < 1e-12


def test_rasterizer_golden_bit_exact(dib, golden_dir):
    bf, ops = dib
    g = _load(golden_dir, "psf_cases.npz")
    n = int(g["n"])
    xs = np.stack([g["x_%d" % k] for k in range(n)])
    fr = np.array([g["meta_%d" % k][1] for k in range(n)])
    raw = ops.rasterize_psfs(xs, fr, "cuda", canvas=256, center=False, out_side=256, dtype=torch.float64).cpu().numpy()
    cen, offs = ops.rasterize_psfs(xs, fr, "cuda", canvas=256, center=True, out_side=256, dtype=torch.float64, return_offsets=True)
    cen = cen.cpu().numpy()
    half = ops.rasterize_psfs(xs, fr, "cuda", canvas=256, center=True, out_side=128, dtype=torch.float16).cpu().numpy()
    for k in range(n):
        ref_raw = np.zeros(256 * 256)
        ref_raw[g["raw_idx_%d" % k]] = g["raw_val_%d" % k]
        assert np.array_equal(raw[k].ravel(), ref_raw), k
        ref_cen = np.zeros(256 * 256)
        ref_cen[g["cen_idx_%d" % k]] = g["cen_val_%d" % k]
        assert np.array_equal(cen[k].ravel(), ref_cen), k
        assert np.array_equal(half[k], ref_cen.reshape(256, 256).astype(np.float16)[64:192, 64:192]), k
        ox, oy = po.centroid_offsets(ref_raw.reshape(256, 256))
        assert tuple(offs[k].cpu().numpy()) == (ox, oy)


def test_checksum(dib):
    bf, ops = dib
    t = torch.rand(3, 100, 333).cuda()
    c1 = ops.checksum(t).item()
    c2 = ops.checksum(t.clone()).item()
    assert c1 == c2
    t2 = t.clone()
    t2[1, 50, 100] += 1e-6
    assert ops.checksum(t2).item() != c1
    # shard-combinable: checksum(whole) == checksum(first half) (+) checksum(second half at its element offset)?  The
    # per-shard values are compared rank by rank, so only determinism and sensitivity are required here.
    bits = t.cpu().numpy().view(np.uint32).ravel().astype(np.uint64)
    idx = np.arange(bits.size, dtype=np.uint64)

    def mix64(z):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))

    with np.errstate(over="ignore"):
        want = mix64((idx << np.uint64(32)) ^ (idx >> np.uint64(32)) ^ (bits * np.uint64(0x9E3779B97F4A7C15)) ^ idx).sum(dtype=np.uint64)
    assert np.uint64(c1 & 0xFFFFFFFFFFFFFFFF) == want


def test_mixed_batch_both_kernels_and_pitched_views(dib):
    """One call with images for the tiled kernel, images only the exact-order kernel takes (zero-pad mode) and a
    pass-through entry; inputs and outputs that are views with padded pitches (rows of a zero-padded batch)."""
    bf, ops = dib
    rng = np.random.default_rng(21)
    np.random.seed(21)
    psfs = []
    for frac in (1 / 10, 1 / 5, 1 / 2):
        p16, _ = po.stored_psf(0.001, frac, np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    ts = ops.compact_taps(_cuda(np.stack(psfs)), normalize=True)
    shapes = [(3, 130, 470), (3, 40, 50), (3, 97, 131), (3, 90, 60)]
    imgs = [rng.random(s, dtype=np.float32) for s in shapes]
    big_in = torch.zeros((3, 160, 512)).cuda()
    big_in[:, :130, :470] = _cuda(imgs[0])
    srcs = [big_in[:, :130, :470], _cuda(imgs[1]), _cuda(imgs[2]), _cuda(imgs[3])]
    big_out = torch.zeros((3, 128, 160)).cuda()
    outs = [None, None, big_out[:, :97, :131], None]
    idx = [0, 1, 2, -1]
    l0 = bf.launch_count()
    got = bf.blur_batch(srcs, ts, idx, outs=outs)
    # one launch per tiled kernel in use (masked program: group width 0, dense program: 2 or 4) + one exact-order launch
    kinds = {ts.meta[i].prog_group_w == 0 for i in (0, 2)}
    assert bf.launch_count() - l0 == len(kinds) + 1
    for k in range(3):
        want = bo.manual_blur(imgs[k], bo.normalize_psf(psfs[idx[k]]))
        g = got[k].cpu().numpy()
        if shapes[k][1] < 64:
            assert np.array_equal(g, want)         # zero-pad mode runs on the exact-order kernel
        else:
            assert np.abs(g - want).max() <= TOL_FP32, k
    assert torch.equal(got[3], srcs[3])            # psf_index -1: passed through
    assert big_out[:, 97:, :].abs().max().item() == 0 and big_out[:, :, 131:].abs().max().item() == 0


def test_many_images_one_call(dib):
    """More images than DIB_MAX_BATCH: the wrapper splits the call; every image still gets its own PSF."""
    bf, ops = dib
    n = 37
    g = torch.Generator().manual_seed(3)
    imgs = [torch.rand((3, 70 + k % 5, 80 + k % 7), generator=g).cuda() for k in range(n)]
    psf = torch.zeros((n, 128, 128)).cuda()
    for k in range(n):
        psf[k, 63 + k % 3, 63 - k % 4] = 1.0       # one-tap PSFs: out = shifted image in the interior
    ts = ops.compact_taps(psf, normalize=True)
    outs = bf.blur_batch(imgs, ts, list(range(n)))
    for k in range(n):
        dy, dx = k % 3, -(k % 4)
        H, W = imgs[k].shape[1:]
        assert torch.equal(outs[k][:, 8:H - 8, 8:W - 8], imgs[k][:, 8 - dy:H - 8 - dy, 8 - dx:W - 8 - dx]), k


def test_half_images_fast_path_tolerance(dib, golden_dir):
    """fp16 images, default path: fp32 accumulation rounded once; the reference's half loop rounds after every tap.
    Stated tolerance 5e-3 against that loop; against the fp32 reference rounded to half it is 1 half-ulp."""
    bf, ops = dib
    g = _load(golden_dir, "blur_cases.npz")
    worst = 0.0
    for n in range(int(g["n"])):
        img16 = _cuda(g["img_%d" % n], torch.float16)
        psfn16 = _cuda(g["psfn_f16_%d" % n])
        ref_half_loop = g["out_f16_%d" % n].astype(np.float64)
        got = bf.manual_blur(img16, psfn16)                          # default: fast path
        assert got.dtype == torch.float16 and tuple(got.shape) == g["out_f16_%d" % n].shape
        err = np.abs(got.cpu().numpy().astype(np.float64) - ref_half_loop).max()
        worst = max(worst, err)
        assert err <= 5e-3, (n, err)
        # the same taps applied in fp32 to the same (half-valued) image, rounded once, is what the fast path must give
        want = bo.manual_blur(g["img_%d" % n].astype(np.float16).astype(np.float32), g["psfn_f16_%d" % n].astype(np.float32))
        assert np.abs(got.cpu().numpy().astype(np.float64) - want.astype(np.float16).astype(np.float64)).max() <= 1e-3, n
    assert worst > 0          # the two really are different computations


@pytest.mark.parametrize("shape", [(3, 200, 300), (1, 65, 449), (2, 130, 1001), (3, 289, 331), (3, 480, 640)])
def test_half_io_inside_the_tiled_kernel(dib, shape):
    """Half images take the tiled kernel directly: rows are widened while they are staged (odd widths give rows at every
    2-byte phase), accumulation is fp32, results are rounded to half once at the store.  Bit-identical to widening with
    torch, running the fp32 tiled kernel and rounding with torch (the fallback path), in reflect and zero-padding mode."""
    bf, ops = dib
    from detectinblur_b200 import _lib
    rng = np.random.default_rng(sum(shape) + 7)
    img = _cuda(rng.random(shape, dtype=np.float32), torch.float16)
    np.random.seed(shape[2])
    psf16, _ = po.stored_psf(0.001, 1 / 2, np.random)
    ts = ops.compact_taps(_cuda(po.crop128(psf16)), normalize=True)          # half PSF, as the engines upload it
    for pad in (None, _lib.PAD_ZERO128):
        before = bf.launch_count()
        fused = bf.blur_batch([img], ts, [0], pad_mode=pad)[0]
        assert bf.launch_count() == before + 1                                # one kernel, no cast passes
        assert fused.dtype == torch.float16 and tuple(fused.shape) == shape and fused.data_ptr() % 16 == 0
        fallback = bf.blur_batch([img], ts, [0], pad_mode=pad, clamp=[False])[0]    # any epilogue request takes the cast path
        assert torch.equal(fused, fallback)
        want = bo.manual_blur(img.float().cpu().numpy(), ts_weights(ts), pad_mode=bo.PAD_ZERO if pad is not None else None)
        err = np.abs(fused.float().cpu().numpy().reshape(want.shape) - want).max()
        assert err <= 1e-3                                                     # one half rounding of values in [0, 1]
    # fused normalize into a half destination, and a caller-provided unaligned destination (falls back, same values)
    mean, std = [[0.485, 0.456, 0.406][:shape[0]]], [[0.229, 0.224, 0.225][:shape[0]]]
    a = bf.blur_batch([img], ts, [0], mean=mean, std=std)[0]
    b = bf.blur_batch([img.float()], ts, [0], mean=mean, std=std)[0]
    assert a.dtype == torch.float16 and (a.float() - b).abs().max().item() <= 4e-3
    odd = torch.empty((shape[0], shape[1], shape[2] + 1), dtype=torch.float16, device="cuda")[:, :, 1:]
    bf.blur_batch([img], ts, [0], outs=[odd])
    assert torch.equal(odd, bf.blur_batch([img], ts, [0])[0])


def test_edge_inputs(dib):
    """Empty batches, nothing to blur, channel counts 1 and 4, a one-tap PSF, an all-zero PSF, and an image far larger
    than the benchmark's (4 x 2100 x 3001: 101 MB, 413 tiles per channel) through size-independent properties."""
    bf, ops = dib
    assert bf.blur_batch([], None, []) == []
    lst = []
    bf.blur_image_list(lst, [], [])
    assert lst == []
    img = torch.rand(3, 70, 90, device="cuda")
    keep = [img]
    bf.blur_image_list(keep, [{"blurring": False}], [torch.zeros(1, device="cuda")])
    assert keep[0] is img                                               # untouched entries keep their identity
    # a shifted delta moves the image (convolution orientation: the image is read at minus the tap offset)
    delta = torch.zeros(128, 128, device="cuda")
    delta[60, 70] = 1.0
    big = torch.rand(4, 2100, 3001, device="cuda")
    out = bf.manual_blur(big, delta)
    assert tuple(out.shape) == (4, 2100, 3001)
    assert torch.equal(out[:, 10:-10, 10:-10], torch.roll(big, shifts=(-3, 7), dims=(1, 2))[:, 10:-10, 10:-10])
    one = bf.manual_blur(torch.rand(1, 300, 500, device="cuda"), delta)  # C == 1 comes back 2-D, as .squeeze() leaves it
    assert tuple(one.shape) == (300, 500)
    # identity PSF: bit-for-bit the input, on both kernels
    ident = torch.zeros(128, 128, device="cuda")
    ident[63, 63] = 1.0
    assert torch.equal(bf.manual_blur(big, ident), big) and torch.equal(bf.manual_blur(big[:, :200, :300], ident, exact=True), big[:, :200, :300])
    # an all-zero PSF: the reference divides by a zero sum and blurs with 16384 NaN taps (blur_functions.py:98, :63); so
    # does this path -- the tap list grows to hold them and the exact-order kernel walks it
    zts = ops.compact_taps(torch.zeros(128, 128, device="cuda"), normalize=True)
    assert zts.counts == [128 * 128] and zts.max_taps >= 128 * 128
    assert torch.isnan(bf.blur_batch([torch.rand(1, 70, 80, device="cuda")], zts, [0])[0]).all()


def test_overlapped_launches_of_independent_batches(dib):
    """DIB_ALGO_OVERLAP (BlurPlan.run(overlap=True)): back-to-back launches of independent batches that share a tap set may
    overlap tail-to-head (programmatic dependent launch, one scheduler slot per launch in flight); every batch must
    come out exactly as when each launch is ordered after the previous one."""
    bf, ops = dib
    rng = np.random.default_rng(77)
    np.random.seed(77)
    psfs = []
    for frac in (1 / 10, 1 / 5, 1 / 2):
        p16, _ = po.stored_psf(0.005, frac, np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    ts = ops.compact_taps(_cuda(np.stack(psfs)), normalize=True)
    batches = [[_cuda(rng.random((3, 300, 500), dtype=np.float32)) for _ in range(3)] for _ in range(4)]
    want = [[t.clone() for t in bf.blur_batch(b, ts, [0, 1, 2])] for b in batches]
    plans = [bf.prepare_blur(b, ts, [0, 1, 2]) for b in batches]
    for rep in range(30):
        for pl in plans:
            pl.run(overlap=True)
    torch.cuda.synchronize()
    for pl, w in zip(plans, want):
        for got, ref in zip(pl.results, w):
            assert torch.equal(got, ref)
    # a dependent launch right after overlapped ones is ordered as usual
    again = bf.blur_batch(batches[0], ts, [0, 1, 2])
    assert all(torch.equal(a, b) for a, b in zip(again, want[0]))


def test_many_single_chunk_psfs_tiled_vs_exact(dib):
    """Low-exposure PSFs fit one program chunk; for those the kernel takes the chunk record from its parameters (rebuilt on
    the host from the PSF summary) instead of loading it.  30 different PSFs, tiled against the exact-order kernel."""
    bf, ops = dib
    np.random.seed(2024)
    psfs = []
    for k in range(30):
        p16, _ = po.stored_psf([0.005, 0.001, 0.00005][k % 3], [1 / 18, 1 / 10, 1 / 5][k % 3 if k < 15 else (k + 1) % 3], np.random)
        psfs.append(po.crop128(p16).astype(np.float32))
    ts = ops.compact_taps(_cuda(np.stack(psfs)), normalize=True)
    assert sum(1 for m in ts.meta if m.prog_chunks == 1) >= 20
    gen = torch.Generator().manual_seed(9)
    imgs = [torch.rand((3, 100 + 7 * (k % 5), 470 + 11 * (k % 4)), generator=gen).cuda() for k in range(30)]
    fast = bf.blur_batch(imgs, ts, list(range(30)))
    exact = bf.blur_batch(imgs, ts, list(range(30)), exact=True)
    for a, b in zip(fast, exact):
        assert (a - b).abs().max().item() <= TOL_FP32


def test_half_io_full_size(dib):
    """BASELINE's full size in fp16 (3 x 800 x 1333: odd width, rows at every 2-byte phase, border tiles on all sides):
    the in-kernel half path against torch casts around the fp32 kernel, plain and with the fused normalize."""
    bf, ops = dib
    gen = torch.Generator().manual_seed(5)
    imgs = [torch.rand((3, 800, 1333), generator=gen).half().cuda() for _ in range(2)]
    np.random.seed(11)
    psfs = [po.crop128(po.stored_psf(e, f, np.random)[0]) for e, f in ((0.005, 1 / 5), (0.00005, 1))]
    ts = ops.compact_taps(_cuda(np.stack(psfs)), normalize=True)
    fused = bf.blur_batch(imgs, ts, [0, 1])
    casts = bf.blur_batch(imgs, ts, [0, 1], clamp=[False, False])
    assert all(torch.equal(a, b) for a, b in zip(fused, casts))
    mean, std = [[0.485, 0.456, 0.406]] * 2, [[0.229, 0.224, 0.225]] * 2
    n16 = bf.blur_batch(imgs, ts, [0, 1], mean=mean, std=std)
    n32 = bf.blur_batch([i.float() for i in imgs], ts, [0, 1], mean=mean, std=std)
    assert max((a.float() - b).abs().max().item() for a, b in zip(n16, n32)) <= 4e-3


def ts_weights(ts):
    """Dense normalised PSF rebuilt from a tap set (what the oracle's manual_blur takes)."""
    ys, xs, ws = ts.taps(0)
    dense = np.zeros((ts.side, ts.side), dtype=np.float32)
    dense[ys, xs] = ws
    return dense
