"""ctypes binding of libdib.so (the C ABI declared in include/dib.h).

The library is built in-tree by ``make -C detectinblur_b200/csrc`` (or ``__graft_entry__.build()``).  There is no
fallback: a missing library is an ImportError, a failing call raises DibError with the library's message.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DIB_LIB_PATH lets kernel experiments point at an alternative build of the same ABI (never a different backend)
LIB_PATH = os.environ.get("DIB_LIB_PATH") or os.path.join(_HERE, "libdib.so")

DIB_F32, DIB_F16, DIB_F64 = 0, 1, 2
PAD_REFLECT128, PAD_ZERO128, PAD_REPLICATE256 = 0, 1, 2
EPI_NOISE, EPI_CLAMP, EPI_GAMMA, EPI_NORMALIZE, EPI_PHILOX = 1, 2, 4, 8, 16
ALGO_AUTO, ALGO_GENERIC, ALGO_TILED = 0, 1, 2
ALGO_OVERLAP = 0x100
ALGO_DEVICE_PLAN = 0x200


def algo_slot(k):
    return (k & 3) << 12
META_TRUNCATED, META_NO_PROGRAM = 1, 2
MAX_BATCH = 32
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY = -1, -2, -3, -4


class DibError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libdib error %d: %s" % (code, message))
        self.code = code


class Tap(ctypes.Structure):
    _fields_ = [("y", ctypes.c_int16), ("x", ctypes.c_int16), ("w", ctypes.c_float)]


class PsfMeta(ctypes.Structure):
    _fields_ = [
        ("count", ctypes.c_int32),
        ("ymin", ctypes.c_int16), ("ymax", ctypes.c_int16),
        ("xmin", ctypes.c_int16), ("xmax", ctypes.c_int16),
        ("sum", ctypes.c_float),
        ("support", ctypes.c_int32),
        ("prog_chunks", ctypes.c_int32),
        ("prog_steps", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("prog_segs", ctypes.c_int32), ("prog_group_w", ctypes.c_int16), ("prog_shear", ctypes.c_int16),
        ("sy", ctypes.c_double), ("sx", ctypes.c_double),
        ("syy", ctypes.c_double), ("sxx", ctypes.c_double), ("sxy", ctypes.c_double),
    ]


class TapsetLayout(ctypes.Structure):
    _fields_ = [("meta_offset", ctypes.c_size_t), ("taps_offset", ctypes.c_size_t), ("prog_offset", ctypes.c_size_t),
                ("prog_bytes_per_psf", ctypes.c_size_t), ("sched_offset", ctypes.c_size_t), ("total_bytes", ctypes.c_size_t)]


class Image(ctypes.Structure):
    _fields_ = [
        ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("noise", ctypes.c_void_p),
        ("C", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("psf_index", ctypes.c_int32),
        ("src_row_pitch", ctypes.c_int64), ("src_chan_pitch", ctypes.c_int64),
        ("dst_row_pitch", ctypes.c_int64), ("dst_chan_pitch", ctypes.c_int64),
        ("pad_mode", ctypes.c_int32), ("epilogue", ctypes.c_int32),
        ("noise_sd", ctypes.c_float), ("gamma", ctypes.c_float),
        ("mean", ctypes.c_float * 4), ("std", ctypes.c_float * 4),
    ]


class ResizeImage(ctypes.Structure):
    _fields_ = [
        ("src", ctypes.c_void_p), ("dst", ctypes.c_void_p),
        ("C", ctypes.c_int32), ("in_h", ctypes.c_int32), ("in_w", ctypes.c_int32),
        ("out_h", ctypes.c_int32), ("out_w", ctypes.c_int32),
        ("pad_h", ctypes.c_int32), ("pad_w", ctypes.c_int32),
        ("normalize", ctypes.c_int32),
        ("src_row_pitch", ctypes.c_int64), ("src_chan_pitch", ctypes.c_int64),
        ("dst_row_pitch", ctypes.c_int64), ("dst_chan_pitch", ctypes.c_int64),
        ("mean", ctypes.c_float * 4), ("std", ctypes.c_float * 4),
    ]


EXPORTS = ("dib_abi_version", "dib_last_error", "dib_device_info", "dib_tapset_layout_for", "dib_compact_taps",
           "dib_blur_batch", "dib_rasterize_psf", "dib_generate_trajectories", "dib_unpack_psfs", "dib_resize_batch", "dib_checksum", "dib_fp32_probe",
           "dib_u8_to_float", "dib_float_to_u8")


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `make -C detectinblur_b200/csrc` (needs nvcc, sm_100a). "
            "detectinblur_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    lib.dib_abi_version.restype = i32
    lib.dib_last_error.restype = ctypes.c_char_p
    lib.dib_device_info.argtypes = [ctypes.POINTER(i32), ctypes.POINTER(i32)]
    lib.dib_tapset_layout_for.argtypes = [i32, i32, ctypes.POINTER(TapsetLayout)]
    lib.dib_compact_taps.argtypes = [vp, i32, i32, i32, i64, i32, vp, i32, vp]
    lib.dib_blur_batch.argtypes = [ctypes.POINTER(Image), i32, vp, i32, i32, ctypes.POINTER(PsfMeta), i32, i32, u64, u64,
                                   ctypes.POINTER(i32), vp]
    lib.dib_rasterize_psf.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.dib_generate_trajectories.argtypes = [u64, u64, vp, i32, i32, ctypes.c_double, ctypes.c_double, vp, vp, vp, vp]
    lib.dib_unpack_psfs.argtypes = [vp, vp, i32, i32, i32, vp, i32, vp]
    lib.dib_resize_batch.argtypes = [ctypes.POINTER(ResizeImage), i32, i32, ctypes.POINTER(i32), vp]
    lib.dib_checksum.argtypes = [vp, i32, i64, vp, i32, vp]
    lib.dib_fp32_probe.argtypes = [i32, vp, ctypes.POINTER(u64), vp]
    lib.dib_u8_to_float.argtypes = [vp, vp, i32, i64, i32, i64, vp]
    lib.dib_float_to_u8.argtypes = [vp, i32, vp, i64, i32, i64, vp]
    for name in EXPORTS:
        if name not in ("dib_last_error",):
            getattr(lib, name).restype = i32
    if lib.dib_abi_version() != 1:
        raise ImportError("libdib.so ABI version %d, expected 1" % lib.dib_abi_version())
    return lib


lib = _load()


def check(code):
    if code != 0:
        raise DibError(code, lib.dib_last_error().decode("utf-8", "replace"))


def tapset_layout(n_psfs, max_taps):
    lay = TapsetLayout()
    check(lib.dib_tapset_layout_for(int(n_psfs), int(max_taps), ctypes.byref(lay)))
    return lay
