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

# The names `import kessler` offers for this path (/root/reference/kessler/__init__.py:13-18), resolved on first use so
# that importing the package stays cheap (no torch import for the ABI-only users).
_LAZY = {"seed": ("util", "seed"), "ConjunctionDataMessage": ("cdm", "ConjunctionDataMessage"), "CDM": ("cdm", "CDM"),
         "Event": ("cdm", "Event"), "GNSS": ("observation_model", "GNSS"), "Radar": ("observation_model", "Radar")}
_SUBMODULES = ("model", "util", "cdm", "observation_model", "events", "montecarlo", "engine", "tle", "sgp4", "distributed",
               "workloads", "stats", "ppl")


def __getattr__(name):
    import importlib
    if name in _LAZY:
        module, attr = _LAZY[name]
        return getattr(importlib.import_module("." + module, __name__), attr)
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module {!r} has no attribute {!r}".format(__name__, name))
