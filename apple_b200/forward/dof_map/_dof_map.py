from __future__ import annotations

import torch


class DofMap:
    """Free <-> full DOF maps, mirror of ``forward/dof_map/_dof_map.py:10-49``.

    Index tensors are int64 CUDA tensors; masks are per component, flattened row-major."""

    def __init__(self, dim: int, n_points: int, fixed_indices: torch.Tensor, fixed_values: torch.Tensor,
                 free_indices: torch.Tensor):
        self.dim = dim
        self.n_points = n_points
        self.fixed_indices = fixed_indices
        self.fixed_values = fixed_values
        self.free_indices = free_indices

    @property
    def n_fixed(self) -> int:
        return self.fixed_indices.numel()

    @property
    def n_free(self) -> int:
        return self.free_indices.numel()

    @property
    def n_full(self) -> int:
        return self.n_points * self.dim

    def to_free(self, full: torch.Tensor) -> torch.Tensor:
        return full.reshape(-1)[self.free_indices]

    def to_free_grad(self, full: torch.Tensor) -> torch.Tensor:
        return full.reshape(-1)[self.free_indices]

    def to_free_hess_diag(self, full: torch.Tensor) -> torch.Tensor:
        return full.reshape(-1)[self.free_indices]

    def to_full(self, free: torch.Tensor) -> torch.Tensor:
        result = torch.empty(self.n_full, dtype=free.dtype, device=free.device)
        result[self.fixed_indices] = self.fixed_values.to(free.dtype)
        result[self.free_indices] = free
        return result.reshape(self.n_points, self.dim)

    def to_full_grad(self, grad_free: torch.Tensor) -> torch.Tensor:
        result = torch.zeros(self.n_full, dtype=grad_free.dtype, device=grad_free.device)
        result[self.free_indices] = grad_free
        return result.reshape(self.n_points, self.dim)

    def free_mask(self) -> torch.Tensor:
        """(n_points, dim) bool tensor, True where the DOF is free."""
        mask = torch.zeros(self.n_full, dtype=torch.bool, device=self.free_indices.device)
        mask[self.free_indices] = True
        return mask.reshape(self.n_points, self.dim)
