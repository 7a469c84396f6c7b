"""limap_b200 — B200-native line triangulation / line refinement behind the limap operator surface.

The CUDA engine lives in limap_b200/csrc and is reached through the C ABI of include/limap_b200.h
(limap_b200/lib/liblimap_b200.so, loaded with ctypes). There is no CPU fallback: importing the
package works anywhere, using an operator without the built library or without a GPU raises.
"""
from . import config  # noqa: F401

__version__ = "0.1.0"
