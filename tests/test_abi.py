"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/dib.h declares,
and the ctypes mirrors of its structs have the sizes a C compiler gives them."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dib.h")


def _declared_symbols():
    text = open(HEADER).read()
    return re.findall(r"DIB_API\s+[\w\s\*]+?\b(dib_\w+)\s*\(", text)


def test_library_loads_and_exports_every_declared_symbol():
    from detectinblur_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 9
    for name in names:
        assert hasattr(_lib.lib, name), "libdib.so does not export %s" % name
    assert set(names) == set(_lib.EXPORTS)
    assert _lib.lib.dib_abi_version() == 1


def test_header_is_plain_c_and_struct_sizes_match(tmp_path):
    from detectinblur_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dib.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(dib_tap), '
                   'sizeof(dib_psf_meta), sizeof(dib_image), sizeof(dib_tapset_layout));'
                   'printf("%zu\\n", sizeof(dib_resize_image));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_lib.Tap), ctypes.sizeof(_lib.PsfMeta), ctypes.sizeof(_lib.Image),
                     ctypes.sizeof(_lib.TapsetLayout), ctypes.sizeof(_lib.ResizeImage)]


def test_argument_validation_without_a_gpu():
    """Validation happens before any CUDA call, so the error paths are checkable on a CPU-only host."""
    from detectinblur_b200 import _lib
    lay = _lib.TapsetLayout()
    assert _lib.lib.dib_tapset_layout_for(0, 16, ctypes.byref(lay)) == _lib.ERR_INVALID
    assert b"n_psfs" in _lib.lib.dib_last_error()
    assert _lib.lib.dib_tapset_layout_for(4, 1024, ctypes.byref(lay)) == 0
    assert lay.total_bytes > lay.prog_offset > lay.taps_offset > 0
    img = (_lib.Image * 1)()
    rc = _lib.lib.dib_blur_batch(img, 40, None, 0, 0, None, 0, 0, 0, 0, None, None)
    assert rc == _lib.ERR_INVALID and b"n_images" in _lib.lib.dib_last_error()
    rc = _lib.lib.dib_blur_batch(img, 1, None, 0, 0, None, 0, 0, 0, 0, None, None)
    assert rc == _lib.ERR_INVALID and b"NULL" in _lib.lib.dib_last_error()
    rz = (_lib.ResizeImage * 1)()
    assert _lib.lib.dib_resize_batch(rz, 0, 0, None, None) == _lib.ERR_INVALID and b"n_images" in _lib.lib.dib_last_error()
    assert _lib.lib.dib_resize_batch(rz, 1, 0, None, None) == _lib.ERR_INVALID and b"NULL" in _lib.lib.dib_last_error()
    assert _lib.lib.dib_unpack_psfs(None, None, 1, 64, 128, None, 1, None) == _lib.ERR_INVALID
    assert _lib.lib.dib_unpack_psfs(ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 200, 128, ctypes.c_void_p(8), 1, None) == _lib.ERR_INVALID
    assert b"canvas" in _lib.lib.dib_last_error()
    rc = _lib.lib.dib_compact_taps(None, 0, 1, 128, 128 * 128, 1, None, 1024, None)
    assert rc == _lib.ERR_INVALID
    rc = _lib.lib.dib_rasterize_psf(ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 2000, 200, 1, 128, ctypes.c_void_p(8), 1, None,
                                    ctypes.c_void_p(8), None)
    assert rc == _lib.ERR_INVALID and b"power of two" in _lib.lib.dib_last_error()


def test_product_refuses_cpu_tensors():
    import torch
    import detectinblur_b200.blur_functions as bf
    psf = torch.zeros(128, 128)
    psf[63, 63] = 1
    with pytest.raises(RuntimeError, match="CUDA"):
        bf.manual_blur(torch.rand(3, 80, 80), psf)
    with pytest.raises(RuntimeError, match="CUDA"):
        bf.blur_image_list([torch.rand(3, 80, 80)], [{"blurring": True}], [psf])
