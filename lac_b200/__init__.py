"""lac_b200 -- B200-native arithmetic-coding hot path of pramasoul/lac.

Python host code over a thin C ABI (include/lac_b200.h, lac_b200/_lib/liblac_b200.so) whose
implementation is hand-written sm_100a CUDA (lac_b200/csrc).  No CPU fallback exists.
"""
from ._ffi import LacError, device_info  # noqa: F401

__all__ = ["LacError", "device_info"]
