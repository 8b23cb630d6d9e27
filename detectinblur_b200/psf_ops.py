"""Device-side PSF operations over libdib.so: tap compaction, PSF rasterisation, PSF-derived metadata, checksums.

Reference code these stand in for:
  TapSet / compact_taps   psf/psf.sum() + psf.nonzero() + per-tap index reads   models/blur_functions.py:63-67,98
  tap_extents             expand_targets' min/max of the nonzero coordinates     utils.py:372-380
  psf_pca                 PCA of the PSF support (theta, lambda scale factors)   transforms.py:366-385
  rasterize_psfs          PSF.fit + centerPSF + crop + float16 cast              motion_blur/generate_PSF.py:31-123,
                                                                                 transforms.py:334-335
  checksum                per-shard result fingerprint (all-gathered like utils.all_gather, utils.py:536-576)
"""
import ctypes
import math
import os

import contextlib

import numpy as np
import torch

from . import _lib

_TORCH_TO_DIB = {torch.float32: _lib.DIB_F32, torch.float16: _lib.DIB_F16, torch.float64: _lib.DIB_F64}


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_NO_SWITCH = contextlib.nullcontext()


def _on_device(device):
    """Context that makes ``device`` current for a C-ABI call; nothing to do (and ~10 us less per call) when it already is."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


def _stream_ptr(device):
    """The caller's current CUDA stream on ``device`` as a raw pointer.  torch.cuda.current_stream() builds a Stream object
    and resolves the device three times (9 us per call, three calls per image in the batch-1 chain); the raw query is the
    same value in 0.3 us."""
    if _raw_stream is not None:
        idx = device.index
        return ctypes.c_void_p(_raw_stream(idx if idx is not None else torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("detectinblur_b200: %s must be a CUDA tensor (there is no CPU path)" % what)


class TapSet(object):
    """Compacted, normalised taps of a batch of PSFs (device buffer) plus a host copy of the per-PSF summary."""

    def __init__(self, buffer, layout, n_psfs, max_taps, side, meta):
        self.buffer = buffer            # uint8 CUDA tensor, layout in include/dib.h
        self.layout = layout
        self.n_psfs = n_psfs
        self.max_taps = max_taps
        self.side = side
        self.meta = meta                # ctypes array of PsfMeta (host)

    @property
    def counts(self):
        return [m.count for m in self._host_meta()]

    def _host_meta(self):
        """The per-PSF summaries on the host; a tap set compacted with ``sync=False`` fetches them on first use."""
        if self.meta is None:
            n = self.n_psfs
            raw = self.buffer[self.layout.meta_offset:self.layout.meta_offset + n * ctypes.sizeof(_lib.PsfMeta)].cpu().numpy().tobytes()
            self.meta = (_lib.PsfMeta * n).from_buffer_copy(raw)
        return self.meta

    def taps(self, index):
        """(ys, xs, weights) of PSF ``index`` as numpy arrays, in accumulation order (copies device -> host)."""
        n = min(self._host_meta()[index].count, self.max_taps)
        off = self.layout.taps_offset + index * self.max_taps * 8
        raw = self.buffer[off:off + n * 8].cpu().numpy().tobytes()
        rec = np.frombuffer(raw, dtype=np.dtype([("y", "<i2"), ("x", "<i2"), ("w", "<f4")]))
        return rec["y"].astype(np.int32), rec["x"].astype(np.int32), rec["w"].copy()

    def tap_extents(self, index, centre=63):
        """(left, top, right, bottom) tap offsets relative to the centre -- utils.py:375-379."""
        m = self._host_meta()[index]
        return m.xmin - centre, m.ymin - centre, m.xmax - centre, m.ymax - centre

    def psf_pca(self, index):
        """(theta_rad, scale_factor_lambda1, scale_factor_lambda2) from the support moments -- transforms.py:366-385."""
        m = self._host_meta()[index]
        n = float(m.support)
        mean_y, mean_x = m.sy / n, m.sx / n
        var_y = m.syy / n - mean_y * mean_y
        var_x = m.sxx / n - mean_x * mean_x
        cov = m.sxy / n - mean_y * mean_x
        root = math.sqrt(math.pow((var_x - var_y) / 2, 2) + math.pow(cov, 2))
        lam1 = (var_x + var_y) / 2 + root
        lam2 = max((var_x + var_y) / 2 - root, 0.0)

        def sigmoid(v):
            return 1 / (1 + math.exp(-v))

        s1 = 1 - (sigmoid(math.sqrt(lam1) / 10) - 0.5) * 0.6
        s2 = 1 - (sigmoid(math.sqrt(lam2) / 10) - 0.5) * 0.6
        theta = -math.atan2(lam1 - var_x, -cov)
        return theta, s1, s2


def compact_taps(psfs, normalize, max_taps=1024, dense_only=None, sync=True):
    """Compact a batch of dense PSFs ([n, k, k] or [k, k] CUDA tensor, float32 / float16) into a TapSet.

    One launch for the batch, then one small device->host copy of the per-PSF summaries (counts, extents,
    program sizes) that the blur launcher plans with.  ``max_taps`` sizes the per-PSF tap list (what the exact-order
    kernel walks); a PSF with more nonzero cells -- e.g. one widened by ``--dilate_psf`` (transforms.py:338-342) -- makes the
    call repeat itself once with a list that holds the largest count, as the reference simply loops over more taps.
    ``dense_only`` (default: the DIB_DENSE_ONLY environment variable) gives small PSFs the dense sheared program too, i.e.
    routes every image to the TMA-staged tiled kernel -- for measurements; the default picks the faster kernel per PSF.
    ``sync=False`` skips the read-back of the summaries: the TapSet then has ``meta = None`` and ``blur_batch`` plans the launch
    on the device (DIB_ALGO_DEVICE_PLAN) -- nothing between the PSFs' upload and the blurred images waits for the host, so the
    chain can be captured in a CUDA graph.  ``max_taps`` must then hold the largest PSF (4096 covers any 128 x 128 motion PSF).
    """
    if dense_only is None:
        dense_only = os.environ.get("DIB_DENSE_ONLY", "0") not in ("0", "", "false", "False")
    _require_cuda(psfs, "psfs")
    if psfs.dim() == 2:
        psfs = psfs.unsqueeze(0)
    if psfs.dim() != 3 or psfs.shape[1] != psfs.shape[2]:
        raise ValueError("psfs must be [n, k, k], got %s" % (tuple(psfs.shape),))
    if psfs.dtype not in (torch.float32, torch.float16):
        raise TypeError("PSF dtype must be float32 or float16, got %s" % psfs.dtype)
    psfs = psfs.contiguous()
    n, side = int(psfs.shape[0]), int(psfs.shape[1])
    lay = _lib.tapset_layout(n, max_taps)
    buf = torch.empty(lay.total_bytes, dtype=torch.uint8, device=psfs.device)
    with _on_device(psfs.device):
        _lib.check(_lib.lib.dib_compact_taps(ctypes.c_void_p(psfs.data_ptr()), _TORCH_TO_DIB[psfs.dtype], n, side,
                                             side * side, (1 if normalize else 0) | (2 if dense_only else 0) | (4 if os.environ.get("DIB_MASKED_ONLY", "0") not in ("0", "") else 0), ctypes.c_void_p(buf.data_ptr()),
                                             int(max_taps), _stream_ptr(psfs.device)))
        if not sync:
            return TapSet(buf, lay, n, int(max_taps), side, None)
        meta_bytes = buf[lay.meta_offset:lay.meta_offset + n * ctypes.sizeof(_lib.PsfMeta)].cpu().numpy().tobytes()
    meta = (_lib.PsfMeta * n).from_buffer_copy(meta_bytes)
    worst = max(m.count for m in meta)
    if worst > max_taps:
        if worst > side * side:
            raise _lib.DibError(_lib.ERR_CAPACITY, "PSF tap count %d exceeds the PSF's %d cells" % (worst, side * side))
        return compact_taps(psfs, normalize, max_taps=1 << (worst - 1).bit_length(), dense_only=dense_only)
    return TapSet(buf, lay, n, int(max_taps), side, meta)


def rasterize_psfs(trajectories, fractions, device, canvas=256, center=True, out_side=128, dtype=torch.float16,
                   return_offsets=False):
    """Rasterise trajectories ([n, iters] complex128, as ``Trajectory.x``) into PSFs on ``device``.

    Bit-identical to ``PSF(canvas, trajectory, [fraction]).fit()`` (+ ``centerPSF()`` + central crop + cast).
    """
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("detectinblur_b200: PSF rasterisation runs on a CUDA device (there is no CPU path)")
    if isinstance(trajectories, torch.Tensor):
        # trajectories already on the device (generate_trajectories): no staging
        if trajectories.dtype != torch.complex128 or not trajectories.is_cuda:
            raise TypeError("device trajectories must be a complex128 CUDA tensor")
        t = trajectories if trajectories.dim() == 2 else trajectories[None]
        t = t.contiguous()
        n, iters = int(t.shape[0]), int(t.shape[1])
        t_traj = torch.view_as_real(t).reshape(-1)
        if isinstance(fractions, torch.Tensor) and fractions.is_cuda:
            t_fr = fractions.to(torch.float64).reshape(-1).expand(n).contiguous()      # already on the device: no staging
        else:
            fr = np.array(np.broadcast_to(np.asarray(fractions, dtype=np.float64), (n,)))
            t_fr = torch.from_numpy(fr).pin_memory().to(device, non_blocking=True)
    else:
        traj = np.ascontiguousarray(np.asarray(trajectories, dtype=np.complex128))
        if traj.ndim == 1:
            traj = traj[None]
        n, iters = traj.shape
        fr = np.array(np.broadcast_to(np.asarray(fractions, dtype=np.float64), (n,)))
        # one pinned staging buffer, one asynchronous upload: [trajectories | fractions] as float64
        stage = torch.empty(n * iters * 2 + n, dtype=torch.float64).pin_memory()
        host = stage.numpy()
        host[:n * iters * 2] = traj.view(np.float64).reshape(-1)
        host[n * iters * 2:] = fr
        dev_in = stage.to(device, non_blocking=True)
        t_traj, t_fr = dev_in[:n * iters * 2], dev_in[n * iters * 2:]
    out = torch.empty((n, out_side, out_side), dtype=dtype, device=device)
    offs = torch.empty((n, 2), dtype=torch.int32, device=device)
    scratch = torch.empty((n, canvas, canvas), dtype=torch.float64, device=device)
    with _on_device(device):
        _lib.check(_lib.lib.dib_rasterize_psf(ctypes.c_void_p(t_traj.data_ptr()), ctypes.c_void_p(t_fr.data_ptr()), n, iters,
                                              int(canvas), 1 if center else 0, int(out_side), ctypes.c_void_p(out.data_ptr()),
                                              _TORCH_TO_DIB[dtype], ctypes.c_void_p(offs.data_ptr()),
                                              ctypes.c_void_p(scratch.data_ptr()), _stream_ptr(device)))
    if return_offsets:
        return out, offs
    return out


def generate_trajectories(n, expl, seed, device, first_index=0, canvas=256, iters=2000, max_len=96, return_big_count=False,
                          indices=None):
    """Draw ``n`` camera-shake trajectories on ``device`` (``Trajectory(canvas, iters, max_len, expl).fit().x`` for each,
    motion_blur/generate_trajectory.py:38-98) from the counter-based generator: trajectory k is a function of
    (seed, first_index + k) only -- or of (seed, indices[k]) when explicit 63-bit ``indices`` are given.  ``expl`` is one
    value or one per trajectory.  Returns a complex128 [n, iters] CUDA tensor that ``rasterize_psfs`` takes as is (no
    host round trip)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("detectinblur_b200: trajectories are generated on a CUDA device here; "
                           "motion_blur.generate_trajectory.Trajectory is the seeded host path")
    ex = torch.as_tensor(np.array(np.broadcast_to(np.asarray(expl, dtype=np.float64), (n,))), device=device)
    out = torch.empty((n, iters), dtype=torch.complex128, device=device)
    big = torch.empty(n, dtype=torch.int32, device=device)
    idx = None
    if indices is not None:
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64).reshape(n), device=device)     # bit pattern of the uint64s
    with _on_device(device):
        _lib.check(_lib.lib.dib_generate_trajectories(int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_index),
                                                      ctypes.c_void_p(idx.data_ptr()) if idx is not None else None, int(n), int(iters),
                                                      float(max_len), float(canvas), ctypes.c_void_p(ex.data_ptr()),
                                                      ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(big.data_ptr()),
                                                      _stream_ptr(device)))
    if return_big_count:
        return out, big
    return out


def unpack_psfs(packed_taps, offsets, device, crop_lo=64, out_side=128, dtype=torch.float16):
    """Expand packed sparse PSFs (psf_bank pack words, one uint32 per tap) into dense [n, out_side, out_side] PSFs on
    ``device``: what transforms.py:301-309 + engine.py:84 deliver, from one pinned upload of the taps only."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("detectinblur_b200: PSFs are unpacked on a CUDA device (there is no CPU path)")
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    packed_taps = np.ascontiguousarray(packed_taps, dtype=np.uint32)
    n = len(offsets) - 1
    if n <= 0 or offsets[0] != 0 or offsets[-1] != len(packed_taps) or np.any(np.diff(offsets) < 0):
        raise ValueError("offsets must be n + 1 non-decreasing tap offsets covering packed_taps")
    if dtype not in _TORCH_TO_DIB:
        raise TypeError("unpack_psfs writes float16/32/64 PSFs")
    # one staging buffer, one host-to-device copy: [offsets | taps] (an empty tap list still needs a valid pointer)
    stage = torch.empty(8 * (n + 1) + 4 * max(len(packed_taps), 1), dtype=torch.uint8).pin_memory()
    host = stage.numpy()
    host[:8 * (n + 1)] = offsets.view(np.uint8)
    host[8 * (n + 1):8 * (n + 1) + 4 * len(packed_taps)] = packed_taps.view(np.uint8)
    dev_buf = stage.to(device, non_blocking=True)
    out = torch.empty((n, out_side, out_side), dtype=dtype, device=device)
    with _on_device(device):
        _lib.check(_lib.lib.dib_unpack_psfs(ctypes.c_void_p(dev_buf.data_ptr() + 8 * (n + 1)), ctypes.c_void_p(dev_buf.data_ptr()),
                                            n, int(crop_lo), int(out_side), ctypes.c_void_p(out.data_ptr()),
                                            _TORCH_TO_DIB[dtype], _stream_ptr(device)))
    dev_buf.record_stream(torch.cuda.current_stream(device))
    return out


def checksum(tensor, out=None, accumulate=False):
    """Order-independent 64-bit checksum of a CUDA tensor's raw element bits -> int64 CUDA tensor of one element."""
    _require_cuda(tensor, "tensor")
    if tensor.dtype not in _TORCH_TO_DIB:
        raise TypeError("checksum supports float16/32/64 tensors")
    t = tensor.contiguous()
    if out is None:
        out = torch.zeros(1, dtype=torch.int64, device=t.device)
        accumulate = False
    with _on_device(t.device):
        _lib.check(_lib.lib.dib_checksum(ctypes.c_void_p(t.data_ptr()), _TORCH_TO_DIB[t.dtype], t.numel(),
                                         ctypes.c_void_p(out.data_ptr()), 1 if accumulate else 0, _stream_ptr(t.device)))
    return out


def u8_to_float(images_u8, dtype=torch.float32, out=None):
    """uint8 CUDA tensor [..., H, W] -> float tensor of the same shape, byte / 255 (torchvision's to_tensor scaling, what
    the reference's DataLoader hands engine.py:80): lets a caller upload bytes -- a quarter of the float32 traffic."""
    _require_cuda(images_u8, "images_u8")
    if images_u8.dtype != torch.uint8:
        raise TypeError("u8_to_float takes uint8 tensors")
    src = images_u8.contiguous()
    W = int(src.shape[-1])
    rows = src.numel() // W
    if out is None:
        out = torch.empty(src.shape, dtype=dtype, device=src.device)
    with _on_device(src.device):
        _lib.check(_lib.lib.dib_u8_to_float(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()), _TORCH_TO_DIB[out.dtype],
                                            rows, W, out.stride(-2) if out.dim() > 1 else W, _stream_ptr(src.device)))
    return out


def float_to_u8(image, out=None):
    """float CUDA tensor [C, H, W] (rows may be pitched) -> dense uint8 tensor: 255 * x clipped to [0, 255] and truncated, the
    uint8 image the --cpu_blur path returns (motion_blur/blur_image.py:147)."""
    _require_cuda(image, "image")
    if image.dtype not in (torch.float32, torch.float16) or image.dim() != 3 or image.stride(2) != 1:
        raise TypeError("float_to_u8 takes float32 / float16 [C, H, W] tensors with contiguous rows")
    C, H, W = (int(v) for v in image.shape)
    if C > 1 and image.stride(0) != image.stride(1) * H:
        image = image.contiguous()
    if out is None:
        out = torch.empty((C, H, W), dtype=torch.uint8, device=image.device)
    with _on_device(image.device):
        _lib.check(_lib.lib.dib_float_to_u8(ctypes.c_void_p(image.data_ptr()), _TORCH_TO_DIB[image.dtype], ctypes.c_void_p(out.data_ptr()),
                                            C * H, W, image.stride(1), _stream_ptr(image.device)))
    return out
