"""msmpscu_b200 -- B200-native implementation of the MDPSCU tabulated EAM/FS hot path.

Layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/mdpscu_b200.h, built
in-tree into libmdpscu_b200.so) and a thin host-side mirror of the reference interface
(force class, neighbour list, integrator, box/control file formats).  No CPU fallback."""
from . import capi, constants  # noqa: F401

__all__ = ["capi", "constants"]
