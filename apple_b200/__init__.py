"""apple_b200: B200-native FEM-elasticity hot path behind liblaf/apple's operator API.

Module layout mirrors ``liblaf.apple`` for the path that is replaced:
``common``, ``warp.fem`` / ``warp.model`` / ``warp.potential`` (operators), ``forward`` (model,
problem, forward solve) and ``optim`` (the PNCG the reference takes from ``liblaf.peach``).
All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/apple_b200.h``; there is no CPU, Warp, Triton or torch.compile fallback.
"""

from . import _lib, common, config, fem, forward, mesh, optim, warp
from ._lib import NativeError

__version__ = "0.1.0"
__all__ = ["NativeError", "common", "config", "fem", "forward", "mesh", "optim", "warp"]
