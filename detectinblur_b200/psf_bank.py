"""Stored-PSF bank: writer, reader and a packed sparse form of it.

Writer -- mirror of the reference's ``dataset_utils/generate_PSFs.py`` (:16-60): same directory layout and file format
(``<destination>psfs/P{1..3}E{0..4}/I{index:06d}``, ``np.save`` of a float16 256 x 256 canvas, no extension), same
seeding (``1337 * worker_index`` for numpy and python RNG) and the same draw order (param-major, then exposure, then
index; ``Trajectory.fit().fit()`` per PSF), so a worker writes byte-identical files.  The trajectories are drawn on the
host; rasterisation, centring and the float16 cast run on the GPU in batches.

Packed form -- a stored PSF holds 13 .. ~270 nonzero cells out of 65 536, yet the reference reads a 131 200-byte file
per image (transforms.py:301-309) and uploads a dense 128 x 128 tensor per image (engine.py:84).  A pack keeps one
32-bit word per nonzero cell of the 256 x 256 canvas, ``y | x << 8 | fp16_bits << 16`` in row-major order, for a whole
``P{p}E{e}`` folder (or one generating worker's slice of it):

    offset  size            field
    0       8               magic  b"DIBPACK1"
    8       4 x uint32      canvas (256), first_index, n_psfs, flags (bit 0: an index table follows)
    24      n x uint32      only with flags bit 0: the ascending PSF indices held (else first_index .. first_index + n - 1)
    ...     (n+1) x uint64  tap offsets (PSF k owns words [off[k], off[k+1]))
    ...     T x uint32      packed taps

It lives next to the folder it replaces as ``<stored_psf_directory>/P{p}E{e}[.w{worker:03d}].dibpack`` and is read
through ``numpy.memmap`` -- no torch, no CUDA, safe in DataLoader workers.  ``load_stored_psf`` (the reader the
``BlurImage`` mirror uses) prefers a pack when one covers the index and falls back to the reference's dense file, and
returns the same float16 128 x 128 array either way.  ``PackedPsfBank.upload`` moves only the taps of a batch to the
GPU (one pinned copy) and expands them there (``dib_unpack_psfs``).
"""
import glob
import os
import random
import struct

import numpy as np

PARAMS = [0.005, 0.001, 0.00005]                  # dataset_utils/generate_PSFs.py:31
FRACTIONS = [1 / 18, 1 / 10, 1 / 5, 1 / 2, 1]     # :32
PACK_MAGIC = b"DIBPACK1"
_HEADER = struct.Struct("<8s4I")


# ------------------------------------------------------------------------------------------------ packed format
def pack_words(psf):
    """Packed taps of one float16 canvas (up to 256 x 256), row-major nonzero order."""
    psf = np.asarray(psf)
    if psf.dtype != np.float16:
        raise TypeError("the stored bank is float16 (generate_PSFs.py:59); got %s" % psf.dtype)
    if psf.ndim != 2 or max(psf.shape) > 256:
        raise ValueError("a pack addresses cells with 8-bit coordinates: canvas up to 256 x 256")
    ys, xs = np.nonzero(psf)
    bits = psf[ys, xs].view(np.uint16).astype(np.uint32)
    return ys.astype(np.uint32) | (xs.astype(np.uint32) << 8) | (bits << 16)


def unpack_words(words, canvas=256):
    """Inverse of ``pack_words``: the dense float16 canvas."""
    words = np.asarray(words, dtype=np.uint32)
    psf = np.zeros((canvas, canvas), dtype=np.float16)
    psf[words & 0xff, (words >> 8) & 0xff] = (words >> 16).astype(np.uint16).view(np.float16)
    return psf


def write_pack(path, indices, psfs, canvas=256):
    """Write one pack holding ``psfs`` (an iterable of float16 canvases).  ``indices`` is the first PSF index (the PSFs
    are then consecutive) or the ascending list of the indices held (a partially present folder)."""
    chunks, offsets = [], [0]
    for psf in psfs:
        if np.asarray(psf).shape != (canvas, canvas):
            raise ValueError("every PSF of a pack must be %d x %d" % (canvas, canvas))
        w = pack_words(psf)
        chunks.append(w)
        offsets.append(offsets[-1] + len(w))
    table = None
    if np.ndim(indices) == 0:
        first = int(indices)
    else:
        table = np.asarray(indices, dtype=np.int64)
        if len(table) != len(chunks) or np.any(np.diff(table) <= 0):
            raise ValueError("indices must be ascending, one per PSF")
        first = int(table[0]) if len(table) else 0
        if len(table) and int(table[-1]) - first + 1 == len(table):
            table = None                                   # consecutive after all
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_HEADER.pack(PACK_MAGIC, canvas, first, len(chunks), 0 if table is None else 1))
        if table is not None:
            f.write(table.astype("<u4").tobytes())
        f.write(np.asarray(offsets, dtype="<u8").tobytes())
        if chunks:
            f.write(np.concatenate(chunks).astype("<u4").tobytes())
    os.replace(tmp, path)
    return len(chunks)


class _Pack(object):
    def __init__(self, path):
        with open(path, "rb") as f:
            head = f.read(_HEADER.size)
        if len(head) < _HEADER.size or head[:8] != PACK_MAGIC:
            raise ValueError("%s is not a PSF pack" % path)
        _, canvas, first, n, flags = _HEADER.unpack(head)
        self.path, self.canvas, self.first, self.n = path, canvas, first, n
        pos = _HEADER.size
        self.table = None
        if flags & 1:
            self.table = np.array(np.memmap(path, dtype="<u4", mode="r", offset=pos, shape=(n,))) if n else np.zeros(0, np.uint32)
            pos += 4 * n
        self.offsets = np.memmap(path, dtype="<u8", mode="r", offset=pos, shape=(n + 1,))
        pos += 8 * (n + 1)
        total = int(self.offsets[-1])
        self.words = np.memmap(path, dtype="<u4", mode="r", offset=pos, shape=(total,)) if total else np.zeros(0, dtype=np.uint32)

    def slot(self, index):
        """Position of PSF ``index`` inside this pack, or -1."""
        if self.table is None:
            return index - self.first if self.first <= index < self.first + self.n else -1
        k = int(np.searchsorted(self.table, index))
        return k if k < self.n and int(self.table[k]) == index else -1

    def taps(self, k):
        return self.words[int(self.offsets[k]):int(self.offsets[k + 1])]


class PackedPsfBank(object):
    """Reader over the packs of a stored-PSF directory (``<dir>/P{p}E{e}*.dibpack``)."""

    def __init__(self, stored_psf_directory):
        self.directory = stored_psf_directory
        self._packs = {}            # (param_index, fraction_index) -> list of _Pack

    def _folder_packs(self, param_index, fraction_index):
        key = (param_index, fraction_index)
        if not self._packs.get(key):              # an empty result is not cached: packs may be written later
            stem = os.path.join(self.directory, "P" + str(param_index) + "E" + str(fraction_index))
            self._packs[key] = [_Pack(p) for p in sorted(glob.glob(stem + ".dibpack") + glob.glob(stem + ".w*.dibpack"))]
        return self._packs[key]

    def words(self, param_index, fraction_index, psf_index):
        """Packed taps of one stored PSF, or None when no pack covers it."""
        for pk in self._folder_packs(param_index, fraction_index):
            k = pk.slot(psf_index)
            if k >= 0:
                return pk.taps(k)
        return None

    def dense(self, param_index, fraction_index, psf_index):
        """What transforms.py:301-309 yields for this file: float16, cropped to 128 x 128 when the canvas is larger."""
        w = self.words(param_index, fraction_index, psf_index)
        if w is None:
            return None
        canvas = self._folder_packs(param_index, fraction_index)[0].canvas
        psf = unpack_words(w, canvas)
        if psf.shape[0] > 128:
            psf = psf[64:128 + 64, 64:128 + 64]
        return psf

    def upload(self, keys, device, dtype=None):
        """Dense PSFs of a batch on ``device`` -- [n, 128, 128] of ``dtype`` (default float16, as engine.py:84 builds them)
        -- from one pinned upload of the taps.  ``keys`` is a list of (param_index, fraction_index, psf_index)."""
        import torch
        from . import psf_ops
        chunks, offsets = [], [0]
        for key in keys:
            w = self.words(*key)
            if w is None:
                raise KeyError("no pack under %s holds PSF %s" % (self.directory, (key,)))
            chunks.append(np.asarray(w))
            offsets.append(offsets[-1] + len(w))
        canvas = self._folder_packs(keys[0][0], keys[0][1])[0].canvas
        crop_lo, side = (64, 128) if canvas > 128 else (0, canvas)
        taps = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
        return psf_ops.unpack_psfs(taps, np.asarray(offsets, np.int64), device, crop_lo=crop_lo, out_side=side,
                                   dtype=torch.float16 if dtype is None else dtype)


def pack_psf_bank(stored_psf_directory, remove_dense=False):
    """Convert a bank written by the reference (or by ``generate_psf_bank``) in place: one pack per P{p}E{e} folder.
    Returns {folder name: number of PSFs packed}."""
    done = {}
    for folder in sorted(glob.glob(os.path.join(stored_psf_directory, "P[0-9]*E[0-9]*"))):
        if not os.path.isdir(folder):
            continue
        files = sorted((f for f in os.listdir(folder) if f.startswith("I") and f[1:].isdigit()), key=lambda f: int(f[1:]))
        if not files:
            continue
        indices = [int(f[1:]) for f in files]

        def canvases():
            for f in files:
                with open(os.path.join(folder, f), "rb") as fh:
                    yield np.load(fh)
        done[os.path.basename(folder)] = write_pack(folder + ".dibpack", indices, canvases())
        if remove_dense:
            for f in files:
                os.remove(os.path.join(folder, f))
    return done


# ------------------------------------------------------------------------------------------------ writer
def generate_psf_bank(destination_path, worker_index=0, num_workers=12, total_num_psfs=12000, device="cuda", batch=256,
                      dense=True, packed=False):
    """Write this worker's slice of the bank; returns the number of PSFs written.

    ``dense`` writes the reference's files; ``packed`` writes ``P{p}E{e}.w{worker:03d}.dibpack`` next to the folders
    (both may be set)."""
    import torch
    from . import psf_ops
    from .motion_blur.generate_trajectory import Trajectory
    slice_size = int(total_num_psfs / num_workers)
    start_index = slice_size * worker_index
    end_index = start_index + slice_size
    np.random.seed(1337 * worker_index)
    random.seed(1337 * worker_index)
    for p in range(len(PARAMS)):
        for e in range(len(FRACTIONS)):
            os.makedirs(destination_path + "psfs/P" + str(p + 1) + "E" + str(e), exist_ok=True)
    written = 0
    for p, param in enumerate(PARAMS):
        for e, exposure in enumerate(FRACTIONS):
            folder = destination_path + "psfs/P" + str(p + 1) + "E" + str(e)
            kept = []
            for lo in range(start_index, end_index, batch):
                hi = min(lo + batch, end_index)
                traj = np.stack([Trajectory(canvas=256, max_len=96, expl=param).fit().fit().x for _ in range(lo, hi)])
                psfs = psf_ops.rasterize_psfs(traj, [exposure] * (hi - lo), device, canvas=256, center=True, out_side=256,
                                              dtype=torch.float16).cpu().numpy()
                for k, index in enumerate(range(lo, hi)):
                    if dense:
                        with open(folder + "/I" + "{:06d}".format(index), "wb") as f:
                            np.save(f, psfs[k])
                    written += 1
                if packed:
                    kept.append(psfs)
            if packed and kept:
                write_pack(folder + ".w{:03d}.dibpack".format(worker_index), start_index, (q for blk in kept for q in blk))
    return written


# ------------------------------------------------------------------------------------------------ reader
_banks = {}


def bank_for(stored_psf_directory):
    """The (cached) PackedPsfBank of a directory; packs are discovered lazily, folder by folder."""
    bank = _banks.get(stored_psf_directory)
    if bank is None:
        bank = _banks[stored_psf_directory] = PackedPsfBank(stored_psf_directory)
    return bank


def load_stored_psf(stored_psf_directory, param_index, fraction_index, psf_index):
    """The reader side, transforms.py:301-309: the stored PSF, cropped to its central 128 x 128.  A pack covering the
    index is preferred (a ~1 KB slice of a memory-mapped file instead of a 131 KB read); the dense file is the fallback."""
    psf = bank_for(stored_psf_directory).dense(param_index, fraction_index, psf_index)
    if psf is not None:
        return psf
    path = stored_psf_directory + "/P" + str(param_index) + "E" + str(fraction_index) + "/I" + "{:06d}".format(psf_index)
    with open(path, "rb") as f:
        psf = np.load(f)
    if psf.shape[0] > 128:
        psf = psf[64:128 + 64, 64:128 + 64]
    return psf
