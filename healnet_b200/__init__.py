"""healnet_b200 — B200-native (sm_100a) implementation of HEALNet's iterative fusion forward.

Drop-in for `healnet.models.HealNet` / `healnet.models.Attention` (reference healnet/models/__init__.py:1-11):
same constructor, parameters, state_dict and forward contract; the arithmetic runs in
`libhealnet_b200.so` (C ABI: include/healnet_b200.h).
"""
from .model import Attention, FeedForward, HealNet, PreNorm  # noqa: F401
from ._lib import HealNetLibraryError, build_library, load_library  # noqa: F401

__all__ = ["HealNet", "Attention", "PreNorm", "FeedForward", "build_library", "load_library",
           "HealNetLibraryError"]
