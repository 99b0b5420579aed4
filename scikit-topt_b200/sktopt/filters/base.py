"""Density-filter interface shared by the Helmholtz and the neighbour-weighted
filter (same four fields and method names as the reference's ``filters/base.py``,
so that user code constructing or subclassing filters keeps working).

The method that applies a filter is ``forward``; ``run`` is also declared by the
reference but nothing implements or calls it -- it is kept as a
name only.  Concrete filters accept NumPy arrays (copied to the device and back)
or CUDA tensors (kept on the device) in ``forward`` / ``gradient``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


def _missing(obj, what: str):
    owner = obj.__name__ if isinstance(obj, type) else type(obj).__name__
    return NotImplementedError(f"{owner} does not implement {what}()")


@dataclass
class BaseFilter:
    mesh: object                                 # sktopt._fem.Mesh (skfem.Mesh in the reference)
    elements_volume: np.ndarray                  # (n_elements,)
    radius: float
    design_mask: Optional[np.ndarray] = None     # bool (n_elements,), None = all design

    @classmethod
    def from_defaults(cls, mesh, elements_volume: np.ndarray, radius: float = 0.3,
                      design_mask: Optional[np.ndarray] = None) -> "BaseFilter":
        raise _missing(cls, "from_defaults")

    def update_radius(self, radius: float, **args):
        """Called by the optimiser when the radius schedule moves."""
        raise _missing(self, "update_radius")

    def forward(self, rho_element):
        """Filtered element densities."""
        raise _missing(self, "forward")

    def gradient(self, v):
        """Adjoint of ``forward`` applied to element sensitivities ``v``."""
        raise _missing(self, "gradient")

    def run(self, rho_element: np.ndarray) -> np.ndarray:
        raise _missing(self, "run")
