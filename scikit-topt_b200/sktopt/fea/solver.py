"""Legacy function API (reference ``fea/solver.py:33-215``): thin wrappers that
forward to the GPU-backed implementations in ``solver_elastic``."""
from __future__ import annotations

import numpy as np

from sktopt.fea import composer
from sktopt.fea import solver_elastic as _se
from sktopt.fea.solver_elastic import LinearSolverConfig


def _cfg(chosen_solver, rtol=1e-5, maxiter=None) -> LinearSolverConfig:
    if isinstance(chosen_solver, LinearSolverConfig):
        return chosen_solver
    return LinearSolverConfig(solver=chosen_solver, rtol=rtol, maxiter=maxiter,
                              allow_fallback_to_spsolve=(chosen_solver == "auto"))


def compute_compliance_simp_basis(basis, free_dofs, dirichlet_dofs, force, E0,
                                  Emin, p, nu0, rho):
    return _se.compute_compliance_basis(
        basis, free_dofs, dirichlet_dofs, force, E0, Emin, p, nu0, rho,
        elem_func=composer.simp_interpolation)


def solve_u(K_cond, F_cond, chosen_solver="auto", rtol: float = 1e-8,
            maxiter: int = None) -> np.ndarray:
    return _se.solve_u(K_cond, F_cond, chosen_solver=_cfg(chosen_solver, rtol, maxiter))


def compute_compliance_basis(basis, free_dofs, dirichlet_dofs, force, E0, Emin,
                             p, nu0, rho, elem_func=composer.simp_interpolation,
                             solver="auto", rtol: float = 1e-5, maxiter: int = None,
                             timer=None):
    return _se.compute_compliance_basis(
        basis, free_dofs, dirichlet_dofs, force, E0, Emin, p, nu0, rho,
        elem_func=elem_func, solver_config=_cfg(solver, rtol, maxiter), timer=timer)


def compute_compliance_basis_numba(basis, free_dofs, dirichlet_dofs, force, E0,
                                   Emin, p, nu0, rho,
                                   elem_func=composer.simp_interpolation,
                                   solver="auto", rtol: float = 1e-5,
                                   maxiter: int = None, n_joblib: int = 1):
    return compute_compliance_basis(basis, free_dofs, dirichlet_dofs, force, E0,
                                    Emin, p, nu0, rho, elem_func=elem_func,
                                    solver=solver, rtol=rtol, maxiter=maxiter)


def solve_multi_load(basis, free_dofs, dirichlet_dofs, force_list, E0, Emin, p,
                     nu0, rho, u_all, solver="auto",
                     elem_func=composer.simp_interpolation, rtol: float = 1e-5,
                     maxiter: int = None, n_joblib: int = 1, timer=None):
    return _se.solve_multi_load(
        basis, free_dofs, dirichlet_dofs, force_list, E0, Emin, p, nu0, rho,
        u_all, solver_config=_cfg(solver, rtol, maxiter), elem_func=elem_func,
        timer=timer)


def compute_compliance_basis_multi_load(basis, free_dofs, dirichlet_dofs,
                                        force_list, E0, Emin, p, nu0, rho, u_all,
                                        solver="auto",
                                        elem_func=composer.simp_interpolation,
                                        rtol: float = 1e-5, maxiter: int = None,
                                        n_joblib: int = 1, timer=None):
    return _se.compute_compliance_basis_multi_load(
        basis, free_dofs, dirichlet_dofs, force_list, E0, Emin, p, nu0, rho,
        u_all, solver_config=_cfg(solver, rtol, maxiter), elem_func=elem_func,
        timer=timer)


__all__ = [
    "solve_u", "compute_compliance_basis", "compute_compliance_basis_numba",
    "compute_compliance_simp_basis", "solve_multi_load",
    "compute_compliance_basis_multi_load",
]
