"""Mirror of the reference's ``motion_blur`` package: trajectories on the host RNG stream, PSF rasterisation on the GPU."""
from .generate_trajectory import Trajectory  # noqa: F401
from .generate_PSF import PSF  # noqa: F401
