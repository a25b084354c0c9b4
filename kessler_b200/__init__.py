"""kessler_b200 -- B200-native implementation of Kessler's Monte Carlo conjunction
hot path (kesslerlib/kessler, kessler/model.py forward simulation).

Layout: ``csrc/`` CUDA kernels + C ABI (``include/kessler_b200.h``), ``_native`` the
ctypes binding, ``engine`` torch-facing wrappers, and the host-side mirror of the
reference's Python surface: ``tle``/``sgp4`` (the dsgp4 calls Kessler makes),
``util``, ``observation_model``, ``model`` (Conjunction, ConjunctionSimplified,
find_conjunction).  No CPU fallback: compute entry points raise without a GPU.
"""
from . import _native  # noqa: F401
from ._native import KesslerCudaError  # noqa: F401

__version__ = "0.1.0"
