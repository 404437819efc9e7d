"""srfdet_b200: B200-native (sm_100a) point-cloud -> region-feature hot path of SRFDet3D.

Kernels: srfdet_b200/csrc (C ABI in include/srfdet_b200.h).  Host mirror of the reference's
plugin interface: srfdet_b200.plugin.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__version__ = '0.1.0'
