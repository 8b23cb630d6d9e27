"""Mirror of the reference's ``models/blur_functions.py`` over the B200 kernels (libdib.so).

Same call signatures and semantics as /root/reference/models/blur_functions.py:
  * ``manual_blur(image_GPU, psf_GPU, add_noise, noise_level, add_block, add_jpeg_artifact, jpeg_compressor)``  (:11-89)
  * ``blur_image_list(images_GPU, blur_dicts, psfs_GPU, ...)`` mutating the list in place                         (:92-100)
plus ``blur_batch`` -- the batched call both are built on (one tap compaction launch + one or two blur launches
for the whole list instead of O(taps) launches and two host syncs per tap per image).

Differences from the reference, all deliberate:
  * results are new tensors whose rows start 16-byte aligned: for a width that is not a multiple of 4 they are views
    ``[:, :, :W]`` of a slightly wider allocation (the reference returns a view too -- a crop of its padded
    accumulator, blur_functions.py:69); aligned rows let the tiled kernel store whole 16-byte quads without per-row
    edge handling.  Pass ``outs=`` to ``blur_batch`` to write anywhere else;
  * CUDA tensors only: there is no CPU path;
  * fp32 images take the tiled kernel (FMA accumulation, <= 1e-5 from the reference loop; measured ~4e-7);
    ``exact=True`` (or ``DIB_EXACT=1``) forces the exact-order kernel, bit-identical to the reference's loop.
  * fp16 images (what the reference's engines pass) are widened, blurred with fp32 accumulation and rounded to half
    once: within 5e-3 of the reference's half loop, which rounds after every tap; ``exact=True`` reproduces that loop
    bit for bit on the exact-order kernel.  The widening and the rounding happen inside the tiled kernel (half rows are
    expanded while they are staged, results are packed at the store); batches it cannot take (noise / clamp / gamma
    epilogues, unaligned caller-provided destinations, tiny images) go through two torch casts around the fp32 path.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib
from . import psf_ops

_DT = {torch.float32: _lib.DIB_F32, torch.float16: _lib.DIB_F16}
_launch_count = 0     # kernels launched by this module (bench.py reads it for "gpu_launches")


def launch_count():
    return _launch_count


def _exact_default():
    return os.environ.get("DIB_EXACT", "0") not in ("0", "", "false", "False")


def pad_mode_for(psf_side, H, W):
    """Boundary mode manual_blur uses -- blur_functions.py:17 (256 branch), :55-58 (zeros below 64, else reflect).

    Raises RuntimeError where the reference does: torch's reflect padding needs pad (64) < dim."""
    if psf_side > 129:
        return _lib.PAD_REPLICATE256
    if H < 64 or W < 64:
        return _lib.PAD_ZERO128
    if H <= 64 or W <= 64:
        raise RuntimeError("Argument #4: Padding size should be less than the corresponding input dimension, "
                           "but got: padding (63, 64) at dimension 2 of input %s" % ([1, -1, H, W],))
    return _lib.PAD_REFLECT128


class BlurPlan(object):
    """A prepared batched blur: image descriptors built once, ``run()`` only crosses the C ABI (one call per <= 32
    images).  ``blur_batch`` is ``prepare_blur(...).run()``; loops that re-blur the same buffers keep the plan."""

    def __init__(self, descs, n, tapset, dtype, algo, philox_seed, device, results, keep):
        self.descs, self.n, self.tapset, self.dtype, self.algo = descs, n, tapset, dtype, algo
        self.philox_seed, self.device, self.results, self._keep = philox_seed, device, results, keep
        self._launches = ctypes.c_int(0)

    def run(self, overlap=False):
        """Launch.  ``overlap=True`` declares this batch independent of whatever was launched just before it on the
        stream (e.g. the previous batch of a loop over independent batches): the kernel may then start on SMs that launch
        has already vacated instead of waiting for its last tile.  Chunks of one oversized batch always overlap."""
        global _launch_count
        ts = self.tapset
        with psf_ops._on_device(self.device):
            stream = psf_ops._stream_ptr(self.device)
            for lo in range(0, self.n, _lib.MAX_BATCH):
                cnt = min(_lib.MAX_BATCH, self.n - lo)
                sub = ctypes.cast(ctypes.byref(self.descs, lo * ctypes.sizeof(_lib.Image)), ctypes.POINTER(_lib.Image))
                algo = self.algo
                if ts is not None:
                    ts.launch_seq = getattr(ts, "launch_seq", 0) + 1      # overlapping launches rotate scheduler slots
                    algo |= _lib.algo_slot(ts.launch_seq)
                    if overlap or lo > 0:
                        algo |= _lib.ALGO_OVERLAP
                    if ts.meta is None and self.algo == _lib.ALGO_AUTO:
                        algo |= _lib.ALGO_DEVICE_PLAN              # no host copy of the summaries: the device plans
                _lib.check(_lib.lib.dib_blur_batch(sub, cnt, ctypes.c_void_p(ts.buffer.data_ptr()) if ts is not None else None,
                                                   ts.n_psfs if ts is not None else 0, ts.max_taps if ts is not None else 0,
                                                   ts.meta if ts is not None else None, _DT[self.dtype], algo,
                                                   int(self.philox_seed or 0), lo, ctypes.byref(self._launches), stream))
                _launch_count += self._launches.value
        return self.results


def prepare_blur(images, tapset, psf_indices, outs=None, noise=None, noise_sd=None, clamp=None, philox_seed=None,
                 mean=None, std=None, gamma=None, exact=None, pad_mode=None):
    """Build the descriptors of a batched blur of CHW CUDA tensors (same dtype, any sizes) with PSFs of ``tapset``.

    images       list of [C, H, W] tensors (float32 or float16, rows contiguous)
    psf_indices  per image: index into ``tapset`` or -1 (pass the image through the epilogue only)
    outs         optional list of destination tensors ([C, H, W'] views with larger pitches allowed: padded batches)
    noise        optional list of pre-drawn N(0,1) tensors (or None entries), ``noise_sd`` the matching sqrt(var)
    philox_seed  draw the noise in-kernel instead (Philox4x32-10); needs ``noise_sd``
    mean, std    optional per-image (C,) sequences: fused ``(x - mean) / std`` (net_transforms.py:135-139)
    pad_mode     boundary mode for every image instead of manual_blur's own rule (``pad_mode_for``); the Fourier-path
                 mirror (motion_blur/blur_image.py) asks for zeros outside the image at any size
    Returns a BlurPlan; ``plan.run()`` launches and returns the list of output tensors.
    """
    n = len(images)
    exact = _exact_default() if exact is None else exact
    if n == 0:
        return BlurPlan((_lib.Image * 1)(), 0, tapset, torch.float32, _lib.ALGO_AUTO, 0, torch.device("cuda"), [], [])
    if images[0].dtype == torch.float16 and not exact and not _half_tiled_ok(images, tapset, psf_indices, outs, noise, noise_sd,
                                                                             clamp, philox_seed, gamma, pad_mode):
        return _HalfPlan(images, tapset, psf_indices, outs, noise, noise_sd, clamp, philox_seed, mean, std, gamma, pad_mode)
    dev = images[0].device
    dtype = images[0].dtype
    if dtype not in _DT:
        raise TypeError("detectinblur_b200 blurs float32 or float16 images, got %s" % dtype)
    results = []
    descs = (_lib.Image * n)()
    keep = []
    # images of one shape and no caller-provided destinations: ONE allocation for all results (views [k, :, :, :W] of an
    # [n, C, H, W'] buffer with 16-byte-aligned rows), so that a caller can move or convert the whole batch in one go
    shared_out = None
    if outs is None and noise is None and n > 1 and all(im.dim() == 3 and im.shape == images[0].shape for im in images):
        C0, H0, W0 = (int(v) for v in images[0].shape)
        quad0 = 16 // images[0].element_size()
        shared_out = torch.empty((n, C0, H0, (W0 + quad0 - 1) // quad0 * quad0), dtype=dtype, device=dev)
    for k, img in enumerate(images):
        psf_ops._require_cuda(img, "image %d" % k)
        if img.dim() != 3:
            raise ValueError("images must be [C, H, W], got %s" % (tuple(img.shape),))
        if img.dtype != dtype or img.device != dev:
            raise TypeError("all images of a batch must share dtype and device")
        if img.stride(2) != 1 or (img.shape[0] > 1 and img.stride(0) < img.stride(1) * img.shape[1]):
            img = img.contiguous()
        keep.append(img)
        C, H, W = (int(s) for s in img.shape)
        if outs is not None and outs[k] is not None:
            out = outs[k]
            if out.dtype != dtype or out.device != dev or out.dim() != 3 or out.stride(2) != 1:
                raise ValueError("bad destination tensor for image %d" % k)
        elif noise is not None and noise[k] is not None and tuple(noise[k].shape) == (C, H, W):
            out = torch.empty_like(noise[k], dtype=dtype)            # a pre-drawn noise tensor shares the destination's layout
        elif shared_out is not None:
            out = shared_out[k, :, :, :W]
        else:
            quad = 16 // img.element_size()
            out = torch.empty((C, H, (W + quad - 1) // quad * quad), dtype=dtype, device=dev)[:, :, :W]    # 16-byte-aligned rows
        results.append(out)
        d = descs[k]
        d.src, d.dst = img.data_ptr(), out.data_ptr()
        d.C, d.H, d.W = C, H, W
        d.psf_index = int(psf_indices[k])
        d.src_row_pitch, d.src_chan_pitch = img.stride(1), img.stride(0)
        d.dst_row_pitch, d.dst_chan_pitch = out.stride(1), out.stride(0)
        if pad_mode is not None:
            d.pad_mode = int(pad_mode)
        else:
            d.pad_mode = pad_mode_for(tapset.side, H, W) if d.psf_index >= 0 else _lib.PAD_REFLECT128
        epi = 0
        if noise_sd is not None and noise_sd[k] is not None:
            nz = noise[k] if noise is not None else None
            if nz is not None:
                nz = nz.to(dtype).expand(C, H, W) if nz.shape != out.shape else nz
                if nz.stride() != out.stride():
                    raise ValueError("noise tensor %d must have the destination's layout" % k)
                keep.append(nz)
                d.noise = nz.data_ptr()
                epi |= _lib.EPI_NOISE
            elif philox_seed is not None:
                epi |= _lib.EPI_NOISE | _lib.EPI_PHILOX
            d.noise_sd = float(noise_sd[k])
            if clamp is None or clamp[k]:
                epi |= _lib.EPI_CLAMP
        elif clamp is not None and clamp[k]:
            epi |= _lib.EPI_CLAMP
        if gamma is not None and gamma[k] is not None:
            epi |= _lib.EPI_GAMMA
            d.gamma = float(gamma[k])
        if mean is not None and mean[k] is not None:
            epi |= _lib.EPI_NORMALIZE
            for c in range(min(C, 4)):
                d.mean[c] = float(mean[k][c])
                d.std[c] = float(std[k][c])
        d.epilogue = epi
    algo = _lib.ALGO_GENERIC if exact else _lib.ALGO_AUTO
    return BlurPlan(descs, n, tapset, dtype, algo, philox_seed, dev, results, keep)


def _half_tiled_ok(images, tapset, psf_indices, outs, noise, noise_sd, clamp, philox_seed, gamma, pad_mode):
    """True when every image of a half batch can take the tiled kernel's fused half I/O (rows widened while they are
    staged, fp32 accumulation, one rounding at the store): nothing but the normalize epilogue, destination rows 16-byte
    aligned (the wrapper's own allocations are), sides above 64 px, PSFs with a tiled program."""
    if noise is not None or noise_sd is not None or clamp is not None or philox_seed is not None or gamma is not None:
        return False
    if pad_mode is not None and pad_mode not in (_lib.PAD_REFLECT128, _lib.PAD_ZERO128):
        return False
    for k, img in enumerate(images):
        if img.dim() != 3 or img.dtype != torch.float16:
            return False
        if psf_indices[k] < 0:
            return False                                   # pass-through images go through the generic kernel's epilogue
        if img.shape[1] <= 64 or img.shape[2] <= 64 or tapset is None or tapset.side > 129:
            return False
        if tapset.meta is not None:       # (device-planned tap sets: the kernel is chosen on the device -- small PSFs take the
            m = tapset.meta[int(psf_indices[k])]           # in-kernel half path, the others the exact-order half loop)
            if m.count <= 0 or m.prog_chunks <= 0 or (m.flags & _lib.META_NO_PROGRAM) or m.prog_group_w != 0:
                return False                               # in-kernel half I/O is the masked kernel's (small PSFs)
        if outs is not None and outs[k] is not None:
            o = outs[k]
            if o.dtype != torch.float16 or o.dim() != 3 or o.data_ptr() % 16 or o.stride(1) % 8 or (o.shape[0] > 1 and o.stride(0) % 8):
                return False
    return True


class _HalfPlan(object):
    """fp16 images on the fast path: widen to fp32, blur with fp32 accumulation (tiled kernel where eligible), round to
    half ONCE.  More accurate than the reference's half loop, which rounds after every tap (it differs from it by up to
    ~5e-3 on [0,1] images, SURVEY.md section 7 "dtype contract"); ``exact=True`` reproduces the half loop bit for bit.
    Fallback of the fused half I/O of the tiled kernel (``_half_tiled_ok``): the two casts are plain torch element-wise
    copies around the fp32 path."""

    def __init__(self, images, tapset, psf_indices, outs, noise, noise_sd, clamp, philox_seed, mean, std, gamma, pad_mode=None):
        self.images, self.tapset, self.psf_indices, self.outs = images, tapset, psf_indices, outs
        self.kw = dict(noise=None if noise is None else [None if z is None else z.float() for z in noise], noise_sd=noise_sd,
                       clamp=clamp, philox_seed=philox_seed, mean=mean, std=std, gamma=gamma, exact=False, pad_mode=pad_mode)

    def run(self, overlap=False):
        wide = [im.float() for im in self.images]
        res32 = prepare_blur(wide, self.tapset, self.psf_indices, **self.kw).run()
        results = []
        for k, r in enumerate(res32):
            if self.outs is not None and self.outs[k] is not None:
                self.outs[k].copy_(r)
                results.append(self.outs[k])
            else:
                results.append(r.to(torch.float16))
        return results


def blur_batch(images, tapset, psf_indices, **kwargs):
    """Blur a list of CHW CUDA tensors in one call; arguments as ``prepare_blur``.  Returns the output tensors."""
    return prepare_blur(images, tapset, psf_indices, **kwargs).run()


def _draw_effects(add_noise, noise_level, add_block, add_jpeg_artifact):
    """Consume numpy's global RNG exactly as one manual_blur call does (blur_functions.py:72-87): the draws do not
    depend on pixel data, so they are made up front, image by image, and applied after the batched launch."""
    d = {"noise_var": None, "block_scale": None, "jpeg_quality": None}
    if add_noise:
        d["noise_var"] = np.random.uniform(0.00000001, noise_level)          # :73
    if add_block:
        if np.random.uniform(0, 1) > 0.5:                                    # :77
            d["block_scale"] = np.random.uniform(0.6, 1)                      # :79
    if add_jpeg_artifact:
        if np.random.uniform(0, 1) > 0.35:                                   # :85
            d["jpeg_quality"] = np.random.uniform(20, 90)                     # :86
    return d


def _post_effects(output, draws, jpeg_compressor):
    """Block and JPEG artefacts stay torch code, exactly as in the reference (blur_functions.py:76-87)."""
    if draws["block_scale"] is not None:
        original_shape = output.shape
        scale_factor = draws["block_scale"]
        output = torch.nn.functional.interpolate(output.unsqueeze(axis=0), scale_factor=(scale_factor, scale_factor),
                                                 mode='nearest').squeeze()
        output = torch.nn.functional.interpolate(output.unsqueeze(axis=0), size=original_shape[1:], mode='nearest').squeeze()
    if draws["jpeg_quality"] is not None:
        from . import transforms
        output = transforms.add_jpeg_artifact_to_image(output, jpeg_compressor, draws["jpeg_quality"])
    return output


def manual_blur(image_GPU, psf_GPU, add_noise=False, noise_level=0.001, add_block=False, add_jpeg_artifact=False,
                jpeg_compressor=None, exact=None):
    """blur_functions.py:11-89.  image is CxHxW, psf is kxk (k <= 129: centre 63; larger: centre 127) and normalised."""
    psf_ops._require_cuda(image_GPU, "image_GPU")
    psf_ops._require_cuda(psf_GPU, "psf_GPU")
    if psf_GPU.dim() != 2:
        raise ValueError("psf must be k x k")
    pad_mode_for(int(psf_GPU.shape[0]), int(image_GPU.shape[1]), int(image_GPU.shape[2]))   # raises like the reference
    # the 0-dim PSF element multiplies in the image dtype (torch type promotion), so taps live in that dtype
    tapset = psf_ops.compact_taps(psf_GPU.to(image_GPU.dtype), normalize=False)
    draws = _draw_effects(add_noise, noise_level, add_block, add_jpeg_artifact)
    noise = noise_sd = None
    if add_noise:
        noise = [torch.randn(image_GPU.shape, dtype=image_GPU.dtype, device=image_GPU.device)]   # :74 randn_like
        noise_sd = [math.sqrt(draws["noise_var"])]
    output = blur_batch([image_GPU], tapset, [0], noise=noise, noise_sd=noise_sd, exact=exact)[0]
    output = output.squeeze()                                                       # :69
    return _post_effects(output, draws, jpeg_compressor)


def blur_image_list(images_GPU, blur_dicts, psfs_GPU, add_noise=False, noise_level=0.001, add_block=False,
                    add_jpeg_artifact=False, jpeg_compressor=None, exact=None, sync=True):
    """blur_functions.py:92-100: blur, in place, every list entry whose ``blur_dict["blurring"]`` is truthy.

    Each PSF is normalised by its own sum (:98) during tap compaction; entries that are not blurred keep their
    identity.  Random draws happen image by image in list order, as in the reference.
    ``sync=False``: the PSF summaries are not read back and the launch is planned on the device, so the
    call returns without waiting for the uploads that feed it (``psf_ops.compact_taps(sync=False)``).
    """
    idx = [k for k, bd in enumerate(blur_dicts) if bd["blurring"]]
    if not idx:
        return None
    # group by (image dtype, PSF side): one compaction + one blur call per group
    groups = {}
    for k in idx:
        psf_ops._require_cuda(images_GPU[k], "images_GPU[%d]" % k)
        psf_ops._require_cuda(psfs_GPU[k], "psfs_GPU[%d]" % k)
        key = (images_GPU[k].dtype, int(psfs_GPU[k].shape[0]), psfs_GPU[k].dtype)
        groups.setdefault(key, []).append(k)
    draws, noise_t = {}, {}
    for k in idx:   # reference order: per blurred image, its numpy draws and its randn_like
        draws[k] = _draw_effects(add_noise, noise_level, add_block, add_jpeg_artifact)
        if add_noise:
            noise_t[k] = torch.randn(images_GPU[k].shape, dtype=images_GPU[k].dtype, device=images_GPU[k].device)
    for (img_dtype, side, psf_dtype), members in groups.items():
        for k in members:
            pad_mode_for(side, int(images_GPU[k].shape[1]), int(images_GPU[k].shape[2]))
        psfs = torch.stack([psfs_GPU[k] for k in members])
        want_exact = _exact_default() if exact is None else exact
        if psf_dtype == img_dtype and not (want_exact and psf_dtype == torch.float32):
            # the compaction kernel's own normalisation: the sum accumulated in fp64 and rounded once, which equals torch's
            # result for every PSF on the fp16 grid summing to at most 1 (all stored / generated PSFs); a general fp32 PSF
            # may differ from torch's reduction tree by one ulp of the sum
            tapset = psf_ops.compact_taps(psfs, normalize=True, **({} if sync else {"sync": False, "max_taps": 4096}))
        elif psf_dtype == img_dtype:
            # exact=True promises the reference's bits for ANY fp32 PSF: normalise with torch itself, PSF by PSF (:98)
            tapset = psf_ops.compact_taps(torch.stack([p / p.sum() for p in psfs]), normalize=False)
        else:
            # the reference multiplies by the 0-dim normalised PSF element cast to the image dtype: normalise in the
            # PSF dtype, cast, then compact, so tap weights carry exactly that rounding
            dense = torch.stack([p / p.sum() for p in psfs])
            tapset = psf_ops.compact_taps(dense.to(img_dtype), normalize=False)
        noise = [noise_t[k] for k in members] if add_noise else None
        noise_sd = [math.sqrt(draws[k]["noise_var"]) for k in members] if add_noise else None
        outs = blur_batch([images_GPU[k] for k in members], tapset, list(range(len(members))), noise=noise,
                          noise_sd=noise_sd, exact=exact)
        for k, out in zip(members, outs):
            images_GPU[k] = _post_effects(out.squeeze(), draws[k], jpeg_compressor)
    return None
