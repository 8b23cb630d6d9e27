"""detectinblur_b200 -- B200 (sm_100a) motion-blur synthesis behind detectInBlur's own Python interface.

Host-side mirror of the reference modules on the blur path (same names, arguments and error behaviour):

    detectinblur_b200.blur_functions      <- models/blur_functions.py      (manual_blur, blur_image_list)
    detectinblur_b200.transforms          <- transforms.py                 (BlurImage)
    detectinblur_b200.motion_blur.*       <- motion_blur/generate_trajectory.py, generate_PSF.py
    detectinblur_b200.net_transforms      <- models/net_transforms.py      (GeneralizedRCNNTransform.normalize hook)

All device work goes through libdib.so (include/dib.h) via ctypes; there is no CPU or torch fallback: importing
``detectinblur_b200._lib`` raises if the library has not been built (``python -c 'import __graft_entry__ as g; g.build()'``).
"""
__all__ = ["blur_functions", "transforms", "motion_blur", "net_transforms", "psf_ops"]
__version__ = "0.1.0"
