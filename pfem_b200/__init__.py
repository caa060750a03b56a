"""pfem_b200: B200-native (sm_100a, fp64) finite-element hot path of PFEM3D behind a C ABI.

The product is `pfem_b200/csrc/libpfem_b200.so` (CUDA kernels + extern "C" entry points
declared in include/pfem_b200.h).  This package holds the host-side mirror of the
reference's Equation/Solver interface for that path; it has no CPU fallback.
"""
from .capi import PfemContext, PfemError, load_library  # noqa: F401
