"""Mirror of the reference's ``motion_blur`` package: trajectories on the host RNG stream, PSF rasterisation and the
``--cpu_blur`` image blur (BlurImageHandler) on the GPU."""
from .generate_trajectory import Trajectory  # noqa: F401
from .generate_PSF import PSF  # noqa: F401
from .blur_image import BlurImageHandler  # noqa: F401
