"""Heat-conduction FEM façade (reference ``fea/solver_heat.py:628-980``).

Filled in after the elasticity path."""
from __future__ import annotations


class FEM_SimpLinearHeatConduction():
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "FEM_SimpLinearHeatConduction: the heat path is not built yet")
