"""Filter constraints (torch parametrizations) -- same classes as `sqfa.constraints`.

These are O(k D) elementwise maps that sit ABOVE the native boundary: the kernels return the
gradient w.r.t. the constrained filters and torch autograd carries it through the constraint to
the raw parameter (reference: /root/reference/src/sqfa/constraints.py).
"""

import torch
import torch.nn as nn

__all__ = ["Sphere", "Identity", "FixedFilters"]


def __dir__():
    return __all__


class Sphere(nn.Module):
    """Keeps every filter (row) on the unit sphere (reference constraints.py:17-54)."""

    def forward(self, X):
        return X / X.norm(dim=-1, keepdim=True)

    def right_inverse(self, S):
        return S


class Identity(nn.Module):
    """No constraint; present so every model has a parametrization (reference constraints.py:58-92)."""

    def forward(self, X):
        return X

    def right_inverse(self, S):
        return S


class FixedFilters(nn.Module):
    """Blocks the gradient of the first `n_row_fixed` filters (reference constraints.py:95-141),
    used by the pairwise training curriculum."""

    def __init__(self, n_row_fixed):
        super().__init__()
        self.n_row_fixed = n_row_fixed

    def forward(self, X):
        frozen = X[: self.n_row_fixed].detach()
        return torch.cat([frozen, X[self.n_row_fixed :]], dim=0)

    def right_inverse(self, X):
        return X
