"""alpha_eigen_b200: B200-native (sm_100a) drop-in for the hot path of alpha-eigen.

Same class surface as the reference package `alpha_eigen_modules` (Slab / Zone / Material /
OneGroup / Phi / OneDirection / TimeUnit / Eigenvalue / inputs); sweeps, source and power
iteration, time stepping and DMD run as hand-written CUDA kernels behind the C ABI declared in
include/alpha_eigen_b200.h.  There is no CPU fallback.
"""
from . import phi as _phi  # noqa: F401  (seeds the global NumPy RNG at import, like phi.py:2)

__all__ = ["slab", "material", "zone", "phi", "eigenvalue", "one_direction", "group", "time_unit", "inputs"]
