"""ripcurrents_b200 -- B200-native (sm_100a) implementation of the per-frame hot path of borgor/ripcurrents:
Farneback dense optical flow, flow aggregation (histograms / thresholds / accumulation / window mean) and
pathline / streakline particle advection, behind the C ABI declared in include/ripcurrents_b200.h.

The product is the CUDA shared library built from ripcurrents_b200/csrc; this package only holds its build script,
a ctypes binding (capi) and the synthetic clip generator used by tests and bench.py.  No CPU fallback exists.
"""
from . import capi  # noqa: F401
from .capi import Context, RcError  # noqa: F401

__all__ = ["capi", "Context", "RcError"]
