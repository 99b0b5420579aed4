"""Filter interface (reference ``filters/base.py:9-37``).  As in the reference
the method that applies the filter is ``forward`` (``run`` is declared there
but never implemented or called)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class BaseFilter():
    mesh: object
    elements_volume: np.ndarray
    radius: float
    design_mask: Optional[np.ndarray] = None

    def update_radius(self, radius: float, **args):
        raise NotImplementedError("")

    @classmethod
    def from_defaults(cls, mesh, elements_volume: np.ndarray, radius: float = 0.3,
                      design_mask: Optional[np.ndarray] = None) -> 'BaseFilter':
        raise NotImplementedError("")

    def run(self, rho_element: np.ndarray) -> np.ndarray:
        raise NotImplementedError("")

    def forward(self, rho_element: np.ndarray) -> np.ndarray:
        raise NotImplementedError("")

    def gradient(self, v: np.ndarray) -> np.ndarray:
        raise NotImplementedError("")
