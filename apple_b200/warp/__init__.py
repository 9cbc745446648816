"""Operator layer: mirror of ``liblaf.apple.warp`` (potentials, model, adapter).

The name is kept so that ``from liblaf.apple.warp.fem import Arap`` becomes
``from apple_b200.warp.fem import Arap``; nothing here uses NVIDIA Warp -- every operator is a
hand-written sm_100a CUDA kernel behind the C ABI of ``include/apple_b200.h``.
"""

from . import fem, model, potential

__all__ = ["fem", "model", "potential"]
