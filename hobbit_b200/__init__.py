"""hobbit_b200 — B200-native (sm_100a) backend for HOBBIT's data-parallel prover hot path.

The product is the CUDA library ``libhobbit_b200.so`` behind the C ABI in ``include/hobbit_b200.h``; this package is a
thin ctypes binding plus the Python mirror of the reference's entry points.  There is no CPU fallback: importing works
anywhere (so the build can be checked), but creating a Context without the built library or without a GPU raises.
"""
from .api import Context, DevF, HobbitError, lib_path, load_library  # noqa: F401

__all__ = ["Context", "DevF", "HobbitError", "lib_path", "load_library"]
