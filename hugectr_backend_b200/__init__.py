"""B200-native Hierarchical Parameter Server lookup path (drop-in for the `hps` Triton backend).

The product is native: ``lib/libhpsx.so`` (engine C ABI, ``include/hpsx.h``) and
``lib/libtriton_hps.so`` (Triton backend C ABI, ``include/triton_hps_backend.h``).  This package is a
thin ctypes binding used by the tests, ``bench.py`` and Python callers; it never computes a lookup
itself and raises if the native library is missing.
"""
from ._native import HpsxError, lib, lib_path  # noqa: F401
from .hps import HPS, LookupSession, ModelParams, SessionStats, ShardGroup, DenseMlp  # noqa: F401

__all__ = ["HPS", "LookupSession", "ModelParams", "SessionStats", "ShardGroup", "DenseMlp", "HpsxError", "lib", "lib_path"]
