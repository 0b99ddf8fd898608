"""term_b200 — B200-native evaluator for term-guard's constraint / analyzer hot path.

Product code is the C-ABI library (csrc/ -> libtermgpu.so, include/termgpu.h); this package is the
Python host mirror of the reference's operator interface used by tests and bench.py.
"""
from ._ffi import TermGpuError, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
