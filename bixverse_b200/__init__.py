"""bixverse_b200: B200-native (sm_100a) drop-in for bixverse's co-expression
correlation hot path.  The package holds only what that path needs: the CUDA
kernels + C ABI (``csrc/`` -> ``libbixcor.so``), the ctypes binding and a
host-side mirror of the reference's ``rs_*`` operator interface."""
from .api import *  # noqa: F401,F403
from . import _lib  # noqa: F401
