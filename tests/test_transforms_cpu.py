"""CPU checks of the BlurImage mirror against the golden blur_dicts made by the unmodified reference: python `random`
consumption, stored-PSF loading, not-blurring defaults, and (through the oracle rasteriser) deferred on-the-fly PSFs."""
import os
import random

import numpy as np
import pytest
from PIL import Image

from oracle import psf_oracle as po

CONFIGS = [
    dict(prob=0.9, use_stored_psfs=False, low_exposure=False, high_exposure=False),
    dict(prob=0.75, blur_type=0.005, use_stored_psfs=False, low_exposure=True),
    dict(prob=1.0, blur_type=0.00005, use_stored_psfs=False, high_exposure=True),
    dict(prob=1.0, blur_type=0.001, blur_exposure=1 / 25, use_stored_psfs=False),
    dict(prob=0.75, blur_type=1, use_stored_psfs=True, low_exposure=True),
    dict(prob=1.0, blur_type=3, use_stored_psfs=True, high_exposure=True),
    dict(prob=0.9, use_stored_psfs=True),
    dict(prob=1.0, use_stored_psfs=False, dont_center_psf=True, blur_type=0.001, low_exposure=True),
    dict(prob=0.0),
]   # must stay in step with tools/make_golden.py:gen_transform_cases


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "transform_cases.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def bank(golden, tmp_path_factory):
    root = tmp_path_factory.mktemp("bank")
    for i, name in enumerate(golden["bank_names"]):
        psf = np.zeros(256 * 256, np.float16)
        psf[golden["bank_idx_%d" % i]] = golden["bank_val_%d" % i]
        path = os.path.join(str(root), str(name))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            np.save(f, psf.reshape(256, 256))
    return str(root)


def test_blur_image_mirror_matches_reference(golden, bank):
    from detectinblur_b200.transforms import BlurImage
    img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (96, 112, 3), dtype=np.uint8))
    for n in range(int(golden["n"])):
        ci, seed = (int(v) for v in golden["cfg_%d" % n])
        kw = dict(CONFIGS[ci])
        kw.setdefault("blur_image_in_transform", False)
        if kw.get("use_stored_psfs"):
            kw["stored_psf_directory"] = bank
        random.seed(seed)
        np.random.seed(seed)
        out_img, _, bd = BlurImage(psf_backend="defer", **kw)(img, None, {})
        assert out_img is img
        assert bool(bd["blurring"]) == bool(golden["blurring_%d" % n]), n
        # python's RNG must have been consumed exactly as the reference consumed it
        assert random.random() == float(golden["next_random_%d" % n]), "case %d: random stream out of step" % n
        want_idx = [None if v == -99 else int(v) for v in golden["indices_%d" % n]]
        got_idx = [None if bd[k] is None else int(bd[k]) for k in ("param_index", "fraction_index")]
        assert got_idx == want_idx, n
        if not bd["blurring"]:
            assert bd["psf"] == [0] and bd["theta_rad"] == 0 and bd["scale_factor_lambda1"] == 1
            continue
        shape = tuple(int(v) for v in golden["psf_shape_%d" % n])
        ref = np.zeros(int(np.prod(shape)), dtype=np.dtype(str(golden["psf_dtype_%d" % n])))
        ref[golden["psf_idx_%d" % n]] = golden["psf_val_%d" % n]
        ref = ref.reshape(shape)
        if bd["psf"] is None:
            # deferred on-the-fly PSF: the mirror drew the trajectory; rasterise it with the oracle and compare
            d = bd["deferred_psf"]
            psf = po.rasterize(d["trajectory"], d["fraction"], 256)
            if d["center"]:
                psf = po.crop128(po.center(psf, 256))
            assert psf.shape == shape
            assert np.array_equal(psf, ref), n
        else:
            assert bd["psf"].dtype == ref.dtype and np.array_equal(bd["psf"], ref), n
            theta, s1, s2 = golden["scalars_%d" % n]
            assert bd["theta_rad"] == theta and bd["scale_factor_lambda1"] == s1 and bd["scale_factor_lambda2"] == s2


def test_packed_bank_reads_like_the_dense_files(golden, bank, tmp_path):
    """The packed sparse bank: same arrays as the reference reader, from ~1 KB per PSF."""
    import shutil
    from detectinblur_b200 import psf_bank
    from detectinblur_b200.transforms import BlurImage
    packed_dir = str(tmp_path / "packed")
    shutil.copytree(bank, packed_dir)
    done = psf_bank.pack_psf_bank(packed_dir, remove_dense=True)
    assert done and all(n >= 1 for n in done.values())
    assert not [f for _, _, fs in os.walk(packed_dir) for f in fs if f.startswith("I")]      # only packs are left
    dense_bytes = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(bank) for f in fs)
    pack_bytes = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(packed_dir) for f in fs)
    assert pack_bytes * 20 < dense_bytes
    checked = 0
    for name in golden["bank_names"]:
        folder, fname = str(name).split("/")[-2:]
        p, e, i = int(folder[1:folder.index("E")]), int(folder[folder.index("E") + 1:]), int(fname[1:])
        want = psf_bank.load_stored_psf(bank, p, e, i)                 # dense file
        got = psf_bank.load_stored_psf(packed_dir, p, e, i)            # pack
        assert got.dtype == np.float16 and got.shape == (128, 128)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16))
        words = psf_bank.bank_for(packed_dir).words(p, e, i)
        assert len(words) == np.count_nonzero(np.load(open(os.path.join(bank, folder, fname), "rb")))
        checked += 1
    assert checked >= 3
    # the BlurImage mirror draws the same blur from the packed bank, RNG stream included
    img = Image.fromarray(np.random.default_rng(3).integers(0, 256, (96, 112, 3), dtype=np.uint8))
    seen = 0
    for n in range(int(golden["n"])):
        ci, seed = (int(v) for v in golden["cfg_%d" % n])
        kw = dict(CONFIGS[ci])
        if not kw.get("use_stored_psfs"):
            continue
        kw.update(blur_image_in_transform=False, stored_psf_directory=packed_dir)
        random.seed(seed)
        np.random.seed(seed)
        _, _, bd = BlurImage(psf_backend="defer", **kw)(img, None, {})
        assert random.random() == float(golden["next_random_%d" % n])
        if bd["blurring"]:
            shape = tuple(int(v) for v in golden["psf_shape_%d" % n])
            ref = np.zeros(int(np.prod(shape)), dtype=np.float16)
            ref[golden["psf_idx_%d" % n]] = golden["psf_val_%d" % n]
            assert np.array_equal(bd["psf"], ref.reshape(shape))
            assert bd["stored_psf_source"][0] == packed_dir
            seen += 1
    assert seen >= 1


def test_pack_format_round_trip_and_worker_slices(tmp_path):
    from detectinblur_b200 import psf_bank
    rng = np.random.default_rng(5)
    canv = []
    for k in range(5):
        c = np.zeros((256, 256), np.float16)
        n = int(rng.integers(0, 40))
        c[rng.integers(0, 256, n), rng.integers(0, 256, n)] = rng.random(n).astype(np.float16)
        canv.append(c)
    canv[2][:] = 0                                            # an empty PSF keeps its slot
    canv[3][0, 0] = canv[3][255, 255] = np.float16(6e-8)       # corners and a subnormal survive
    for c in canv:
        assert np.array_equal(psf_bank.unpack_words(psf_bank.pack_words(c)).view(np.uint16), c.view(np.uint16))
    d = str(tmp_path)
    psf_bank.write_pack(os.path.join(d, "P1E0.w000.dibpack"), 0, canv[:2])
    psf_bank.write_pack(os.path.join(d, "P1E0.w001.dibpack"), 2, canv[2:])
    bank = psf_bank.PackedPsfBank(d)
    for k, c in enumerate(canv):
        assert np.array_equal(bank.dense(1, 0, k), c[64:192, 64:192])
    assert bank.words(1, 0, 5) is None and bank.words(2, 0, 0) is None and bank.dense(1, 0, 7) is None
    with pytest.raises(TypeError):
        psf_bank.pack_words(np.zeros((256, 256), np.float32))
    with open(os.path.join(d, "P2E0.dibpack"), "wb") as f:
        f.write(b"not a pack at all, really not")
    with pytest.raises(ValueError):
        psf_bank.PackedPsfBank(d).words(2, 0, 0)


def test_preblurred_passthrough():
    from detectinblur_b200.transforms import BlurImage
    img = Image.new("RGB", (80, 70))
    out, tgt, bd = BlurImage(prob=1.0, blur_image_in_transform=False, psf_backend="defer")(img, "t", {"preBlurred": True})
    assert out is img and tgt == "t" and bd["blurring"] is False and bd["psf"] == [0] and bd["inverseWarp"] is None


def test_trajectory_mirror_bit_exact(golden_dir):
    from detectinblur_b200.motion_blur.generate_trajectory import Trajectory
    g = np.load(os.path.join(golden_dir, "psf_cases.npz"), allow_pickle=False)
    for k in range(int(g["n"])):
        expl, frac, seed = g["meta_%d" % k]
        np.random.seed(int(seed))
        tr = Trajectory(canvas=256, max_len=96, expl=expl).fit().fit()
        assert np.array_equal(tr.x, g["x_%d" % k]), k
        assert tr.x.dtype == np.complex128 and len(tr.x) == 2000


def test_host_rules_without_a_gpu():
    """Pure host logic of the mirrors: boundary-mode rule, resize scale / padded batch extent, fp16 fast-path gate."""
    import torch
    from detectinblur_b200 import _lib
    from detectinblur_b200 import blur_functions as bf
    from detectinblur_b200 import net_transforms as nt
    from oracle import resize_oracle as ro
    # manual_blur's boundary rule (blur_functions.py:17, :55-58): zeros below 64 px, reflect above, 64 itself raises
    assert bf.pad_mode_for(128, 63, 500) == _lib.PAD_ZERO128 and bf.pad_mode_for(128, 500, 40) == _lib.PAD_ZERO128
    assert bf.pad_mode_for(128, 65, 65) == _lib.PAD_REFLECT128 and bf.pad_mode_for(256, 65, 65) == _lib.PAD_REPLICATE256
    for hw in ((64, 500), (500, 64)):
        with pytest.raises(RuntimeError, match="Padding size should be less"):
            bf.pad_mode_for(128, *hw)
    # resize scale and padded extent (net_transforms.py:36-48, :236-247)
    for h, w in ((480, 640), (640, 427), (800, 1333), (333, 500), (1200, 300)):
        assert nt.resize_scale(h, w, 800, 1333) == ro.resize_scale(h, w, 800, 1333)
    assert nt.padded_batch_shape([(800, 1333), (750, 1000)]) == (800, 1344)
    assert nt.padded_batch_shape([(97, 131)]) == (128, 160)

    # the gate of the in-kernel half path: only the normalize epilogue, aligned destinations, sides > 64, a tiled program
    class Meta(object):
        def __init__(self, count=20, chunks=1, flags=0, group_w=0):
            self.count, self.prog_chunks, self.flags, self.prog_group_w = count, chunks, flags, group_w

    class Ts(object):
        side = 128
        meta = [Meta(), Meta(flags=_lib.META_NO_PROGRAM), Meta(count=0), Meta(count=200, chunks=4, group_w=2)]

    img = torch.zeros((3, 100, 131), dtype=torch.float16)
    ok = lambda **kw: bf._half_tiled_ok(kw.pop("images", [img]), Ts(), kw.pop("idx", [0]), kw.pop("outs", None), kw.pop("noise", None),
                                        kw.pop("noise_sd", None), kw.pop("clamp", None), kw.pop("philox_seed", None),
                                        kw.pop("gamma", None), kw.pop("pad_mode", None))
    assert ok()
    assert ok(pad_mode=_lib.PAD_ZERO128) and not ok(pad_mode=_lib.PAD_REPLICATE256)
    assert not ok(idx=[1]) and not ok(idx=[2]) and not ok(idx=[-1])
    assert not ok(idx=[3])          # a large PSF (dense program): the masked kernel owns the in-kernel half path
    assert not ok(clamp=[False]) and not ok(noise_sd=[0.1]) and not ok(gamma=[2.2]) and not ok(philox_seed=1)
    assert not ok(images=[torch.zeros((3, 64, 131), dtype=torch.float16)])
    assert not ok(images=[img.float()])
    aligned = torch.zeros((3, 100, 136), dtype=torch.float16)[:, :, :131]
    assert ok(outs=[aligned]) and not ok(outs=[torch.zeros((3, 100, 132), dtype=torch.float16)[:, :, 1:]])
    assert not ok(outs=[aligned.float()])
