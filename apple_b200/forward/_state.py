from dataclasses import dataclass

import torch


@dataclass
class ModelState:
    """``forward/_state.py:8-10``: the full displacement field (n_points, 3)."""

    u: torch.Tensor
