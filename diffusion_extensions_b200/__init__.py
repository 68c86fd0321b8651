"""diffusion_extensions_b200 -- B200-native (sm_100a) SO(3) manifold-diffusion hot path.

Drop-in for the reference's ``util`` (SO(3) part), ``distributions.IsotropicGaussianSO3`` and
``diffusion.SO3Diffusion`` / ``ProjectedSO3Diffusion``; hand-written CUDA kernels behind a C ABI
(include/so3d.h, built into libso3d.so by ``python -m diffusion_extensions_b200.build``).
CUDA float32 only -- there is no CPU fallback.
"""
from . import _lib, ops  # noqa: F401
from .ops import manual_seed  # noqa: F401
from . import util, distributions, diffusion  # noqa: F401
from .distributions import IsotropicGaussianSO3, IGSO3xR3, Bingham  # noqa: F401
from .diffusion import SO3Diffusion, ProjectedSO3Diffusion, SE3Diffusion, ProjectedSE3Diffusion  # noqa: F401
from .util import AffineT, AffineGrad  # noqa: F401
from .denoiser import RotPredict, SinusoidalPosEmb  # noqa: F401
from . import parallel  # noqa: F401

__version__ = "0.1.0"
