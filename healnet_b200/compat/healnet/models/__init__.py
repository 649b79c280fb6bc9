"""`healnet.models` of the import shim: the two classes of the hot path (reference healnet/models/__init__.py:1)."""
from healnet_b200 import Attention, HealNet

__all__ = ["HealNet", "Attention"]
