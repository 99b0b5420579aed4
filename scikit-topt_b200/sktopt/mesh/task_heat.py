"""Steady heat-conduction task (reference ``mesh/task_heat.py``)."""
from __future__ import annotations

import re
from dataclasses import dataclass, fields
from typing import List, Literal, Optional, Union

import numpy as np

from sktopt._fem import facet_load, facet_mass
from sktopt.mesh.task_common import FEMDomain

_lit_bc = Literal['u^1', 'u^2', 'u^3', 'all']
_OBJECTIVES = ("compliance", "heat_exchange", "averaged_temp")


def setdiff1d(a, b):
    a = np.asarray(a)
    return np.ascontiguousarray(a[~np.isin(a, b)])


def _per_patch(value, n):
    return [value] * n if isinstance(value, float) else value


def assemble_surface_neumann(basis, neumann_facets_ids, neumann_value):
    """Load vectors of prescribed normal fluxes q_n (``mesh/task_heat.py:23-57``)."""
    as_list = lambda x: x if isinstance(x, list) else [x]
    facets_l, vals_l = as_list(neumann_facets_ids), as_list(neumann_value)
    if len(facets_l) != len(vals_l):
        raise ValueError("Lengths of facets_list and vals_list must match when lists.")
    out = [facet_load(basis.mesh, np.asarray(f, dtype=int), float(q))
           for f, q in zip(facets_l, vals_l)]
    return out[0] if len(out) == 1 else out


def assemble_surface_robin(basis, robin_facets_ids, robin_coefficient,
                           robin_bc_value, rho: Optional[np.ndarray] = None,
                           p: Optional[float] = None):
    """Robin facet matrices  int_G h u v  and loads  int_G h T_env v, one per
    patch (``mesh/task_heat.py:60-139``).  In the reference the density-weighted
    branch is unreachable (``rho_field`` is never set), so the
    matrices do not depend on ``rho``; ``rho``/``p`` are accepted and ignored."""
    facets_l = robin_facets_ids if isinstance(robin_facets_ids, list) else [robin_facets_ids]
    h_l = _per_patch(robin_coefficient, len(facets_l))
    T_l = _per_patch(robin_bc_value, len(facets_l))
    if not (len(facets_l) == len(h_l) == len(T_l)):
        raise ValueError("Lengths of robin_facets_ids and robin_value must match when lists.")
    bilinear, linear = [], []
    for facets, h, Tenv in zip(facets_l, h_l, T_l):
        ids = np.asarray(facets, dtype=int)
        bilinear.append(facet_mass(basis.mesh, ids, float(h)))
        linear.append(facet_load(basis.mesh, ids, float(h) * float(Tenv)))
    return bilinear, linear


@dataclass
class LinearHeatConduction(FEMDomain):
    k: float
    robin_bilinear: Optional[list] = None
    robin_linear: Optional[list] = None
    objective: Literal["compliance", "heat_exchange", "averaged_temp"] = "compliance"
    avg_temp_weight: float = 0.0

    def update_robin_bc(self, rho: np.ndarray, p: float):
        self.robin_bilinear, self.robin_linear = assemble_surface_robin(
            self.basis, robin_facets_ids=self.robin_facets_ids,
            robin_coefficient=self.robin_coefficient,
            robin_bc_value=self.robin_bc_value, rho=rho, p=p,
        )

    @property
    def material_coef(self) -> float:
        return self.k

    @property
    def n_tasks(self) -> int:
        return 1 if isinstance(self.dirichlet_values, float) else len(self.dirichlet_values)

    @classmethod
    def from_facets(cls, basis, dirichlet_facets_ids, dirichlet_values,
                    robin_facets_ids, robin_coefficient, robin_bc_value,
                    design_robin_boundary, design_elements, k: float,
                    objective: str = "compliance",
                    avg_temp_weight: float = 0.0) -> 'LinearHeatConduction':
        if objective not in _OBJECTIVES:
            raise ValueError(
                "objective must be one of 'compliance', 'heat_exchange', or 'averaged_temp'")
        base = FEMDomain.from_facets(
            basis, dirichlet_facets_ids, None, dirichlet_values,
            None, None, None,
            robin_facets_ids, robin_coefficient, robin_bc_value,
            design_robin_boundary, design_elements,
        )
        if robin_facets_ids is not None:
            rb, rl = assemble_surface_robin(
                base.basis, robin_facets_ids=robin_facets_ids,
                robin_coefficient=base.robin_coefficient,
                robin_bc_value=base.robin_bc_value)
        else:
            rb, rl = None, None
        shared = {f.name: getattr(base, f.name) for f in fields(FEMDomain)}
        return cls(**shared, k=k, robin_bilinear=rb, robin_linear=rl,
                   objective=objective, avg_temp_weight=avg_temp_weight)

    @classmethod
    def from_mesh_tags(cls, basis, dirichlet_values, robin_coefficient,
                       robin_bc_value, design_robin_boundary, k: float,
                       objective: str = "compliance",
                       avg_temp_weight: float = 0.0) -> 'FEMDomain':
        mesh = basis.mesh
        design_elements = mesh.subdomains["design"]
        keys = mesh.boundaries.keys()

        def tagged(prefix):
            found = [t for t in keys if re.match(prefix + r"_\d+$", t)]
            found.sort(key=lambda x: int(re.search(r"\d+$", x).group()))
            if found:
                return [mesh.boundaries[t] for t in found]
            return mesh.boundaries[prefix] if prefix in keys else None

        return cls.from_facets(
            basis, tagged("dirichlet"), dirichlet_values, tagged("robin"),
            robin_coefficient, robin_bc_value, design_robin_boundary,
            design_elements, k, objective, avg_temp_weight,
        )
