"""`from lib.pafprocess import pafprocess` (paf_to_pose.py:7 of the reference) resolves to the ctypes-backed module
below, which binds the SWIG-compatible C symbols of libb200pose.so (include/b200pose.h section 4)."""
from . import pafprocess  # noqa: F401
