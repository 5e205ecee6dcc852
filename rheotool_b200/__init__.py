"""rheotool_b200 — B200-native viscoelastic stress step behind rheoTool's constitutiveEq API.

Only the hot path of SURVEY.md §8 lives here: csrc/ (CUDA kernels for sm_100a + the C-ABI + the host
mesh services) and the Python mirror of the reference's plugin interface used by tests and bench.py.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
