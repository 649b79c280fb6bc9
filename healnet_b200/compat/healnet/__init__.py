"""Import shim: `healnet` resolved to the B200 drop-in (put healnet_b200/compat on PYTHONPATH).

Code written against the reference package layout — `from healnet.models import HealNet, Attention`
(reference healnet/models/__init__.py:1-11, healnet/__init__.py:1) — then runs on libhealnet_b200.so unchanged; the
reference's own unit tests (healnet/tests/test_healnet.py) are executed this way by tests/test_gpu_reference_tests.py.
Only the hot-path classes exist here: the reference's losses, baselines and pipeline are out of scope.
"""
from .models import Attention, HealNet  # noqa: F401
