"""Process-wide defaults (the reference's only global switch is the JAX x64 flag that selects the
floating-point type, ``warp/fem/_base.py:100-104``)."""

import torch

from . import _lib

default_dtype: torch.dtype = torch.float32
scatter: int = _lib.SCATTER_TILE  # assembly strategy of the element kernels


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.NativeError("apple_b200 needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())
