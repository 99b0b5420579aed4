"""Elasticity state solves, compliance and element strain energy on the GPU.

Public names follow reference ``fea/solver_elastic.py``: ``LinearSolverConfig``
(:34-40), ``solve_u`` (:61-143), ``compute_compliance_basis`` (:146-237),
``solve_multi_load`` (:240-395), ``compute_compliance_basis_multi_load``
(:398-467), ``strain_energy_skfem_multi`` (:507-537) and the façade
``FEM_SimpLinearElasticity`` (:540-676).

Every solver selector is served by the device-resident Jacobi-PCG
(``csrc/linalg.cu``): there is no direct solver, no PyAMG / PETSc and no CPU
fallback in this build.  The selectors are still validated so configurations
written for the reference construct unchanged.  Arrays may be NumPy (copied to
/ from the device at the API edge, as the reference's in-place ``u_dofs``
contract requires) or CUDA tensors (kept on the device).
"""
from __future__ import annotations

from contextlib import contextmanager
from dataclasses import dataclass
from typing import Callable, Literal

import numpy as np
import torch

from sktopt._b200 import device as dev
from sktopt.fea import composer
from sktopt.fea._engine import KE_ELASTIC, get_engine
from sktopt.fea._petsc_compat import PETScOptions
from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)

ElasticSolver = Literal['cg_jacobi', 'spsolve', 'cg_pyamg', 'petsc', 'petsc_spdirect']
SolverSelector = Literal['auto'] | ElasticSolver
_KNOWN_SOLVERS = ('auto', 'cg_jacobi', 'spsolve', 'cg_pyamg', 'petsc', 'petsc_spdirect')


@dataclass(frozen=True)
class LinearSolverConfig:
    solver: SolverSelector = "spsolve"
    rtol: float = 1e-8
    maxiter: int | None = None
    petsc_options: PETScOptions | None = None
    allow_fallback_to_spsolve: bool = False


def normalize_linear_solver_config(solver, rtol: float = 1e-8, maxiter=None,
                                   petsc_options=None,
                                   allow_fallback_to_spsolve: bool = False
                                   ) -> LinearSolverConfig:
    if isinstance(solver, LinearSolverConfig):
        return solver
    return LinearSolverConfig(solver=solver, rtol=rtol, maxiter=maxiter,
                              petsc_options=petsc_options,
                              allow_fallback_to_spsolve=allow_fallback_to_spsolve)


def _check_solver(name):
    if name not in _KNOWN_SOLVERS:
        raise ValueError(f"Unknown solver: {name}")


@contextmanager
def _noop():
    yield


def _section(timer, name):
    return timer.section(name) if timer else _noop()


def _is_dev(a) -> bool:
    return isinstance(a, torch.Tensor) and a.is_cuda


def solve_u(K_cond, F_cond, chosen_solver='spsolve', rtol: float = 1e-8,
            maxiter: int = None, petsc_options=None,
            allow_fallback_to_spsolve: bool = False) -> np.ndarray:
    """Solve an already enforced SciPy system on the GPU with Jacobi-PCG."""
    cfg = normalize_linear_solver_config(chosen_solver, rtol=rtol, maxiter=maxiter,
                                         petsc_options=petsc_options)
    _check_solver(cfg.solver)
    from sktopt.fea._engine import default_maxiter
    K = K_cond.tocsr()
    K.sort_indices()
    n = K.shape[0]
    rp, ci, va = dev.to_dev(K.indptr, dev.I32), dev.to_dev(K.indices, dev.I32), dev.to_dev(K.data)
    b = dev.to_dev(F_cond)
    x = torch.zeros(n, dtype=dev.F64, device="cuda")
    minv = dev.csr_inv_diag(rp, ci, va)
    pcg = dev.PcgSolver(n)
    dpn = 3 if K.nnz > 40 * n else 1
    pcg.solve(rp, ci, va, minv, b, x, dpn_hint=dpn, rtol=cfg.rtol,
              maxiter=default_maxiter(n) if cfg.maxiter is None else cfg.maxiter)
    logger.info(f"PCG (Jacobi, device) iterations: {pcg.last_iters}, "
                f"converged: {pcg.last_converged}")
    return x.cpu().numpy()


def _run_loads(basis, dirichlet_dofs, force_list, E0, Emin, p, nu0, rho, u_all,
               elem_func, solver_cfg, timer):
    """assemble once -> enforce -> one PCG per load; fills u_all in place and
    returns (compliances, engine)."""
    _check_solver(solver_cfg.solver)
    eng = get_engine(basis, dirichlet_dofs, KE_ELASTIC, nu0)
    # 'cg_jacobi' asks for the diagonal preconditioner; every other selector
    # gets the multigrid V-cycle when the mesh is an eligible tensor grid
    eng.mg_enabled = solver_cfg.solver != "cg_jacobi"
    rho_d = dev.to_dev(rho)
    with _section(timer, "assemble"):
        eng.set_modulus(rho_d, E0, Emin, p, ramp=composer.is_ramp(elem_func))
    with _section(timer, "enforce_bc"):
        # assembled path: the Dirichlet mask is applied while the values are
        # gathered; matrix-free path: only the diagonal + coarse operators
        eng.prepare()
    n_loads = len(force_list)
    comp = np.empty(n_loads)
    with _section(timer, "solve"):
        rhs = []
        for i, f in enumerate(force_list):
            f_d = f if _is_dev(f) else dev.to_dev(f)
            b = eng.rhs if n_loads == 1 else eng.rhs_slot(i)
            dev.enforce_rhs(f_d, None, eng.dir_mask, None, out=b)
            rhs.append(b)
        us = eng.solve_many(rhs, solver_cfg.rtol, solver_cfg.maxiter)
        for i, u in enumerate(us):
            comp[i] = dev.dot(rhs[i], u)
            if u_all is not None:
                if _is_dev(u_all):
                    (u_all[:, i] if u_all.ndim == 2 else u_all).copy_(u)
                else:
                    u_all[:, i] = u.cpu().numpy()
    return comp, eng


def compute_compliance_basis(basis, free_dofs, dirichlet_dofs, force, E0, Emin,
                             p, nu0, rho,
                             elem_func: Callable = composer.simp_interpolation,
                             solver_config: LinearSolverConfig | None = None,
                             solver='auto', rtol: float = 1e-5, maxiter=None,
                             petsc_options=None, timer=None) -> tuple:
    """Single load: returns (compliance, u) with u a NumPy vector."""
    cfg = solver_config if solver_config is not None else \
        normalize_linear_solver_config(solver, rtol=rtol, maxiter=maxiter,
                                       petsc_options=petsc_options)
    comp, eng = _run_loads(basis, dirichlet_dofs, [force], E0, Emin, p, nu0, rho,
                           None, elem_func, cfg, timer)
    return float(comp[0]), eng.solution(0).cpu().numpy()


def solve_multi_load(basis, free_dofs, dirichlet_dofs, force_list, E0, Emin, p,
                     nu0, rho, u_all, solver='auto', solver_config=None,
                     elem_func: Callable = composer.simp_interpolation,
                     rtol: float = 1e-5, maxiter=None, petsc_options=None,
                     timer=None) -> np.ndarray:
    """Shared-stiffness multi-load solve; fills ``u_all`` (n_dof, n_loads).

    Like the reference this returns the compliance array for a single load and
    the stack of enforced right-hand sides otherwise.  The reference only
    offers LU / PETSc for several loads (:334-340); here each load is one PCG
    solve on the once-assembled matrix."""
    cfg = solver_config if solver_config is not None else \
        normalize_linear_solver_config(solver, rtol=rtol, maxiter=maxiter,
                                       petsc_options=petsc_options)
    comp, eng = _run_loads(basis, dirichlet_dofs, force_list, E0, Emin, p, nu0,
                           rho, u_all, elem_func, cfg, timer)
    if len(force_list) == 1:
        return comp
    mask = eng.dir_mask.cpu().numpy().astype(bool)
    cols = []
    for f in force_list:
        fe = np.array(f.cpu().numpy() if _is_dev(f) else f, dtype=np.float64, copy=True)
        fe[mask] = 0.0
        cols.append(fe)
    return np.column_stack(cols)


def compute_compliance_basis_multi_load(basis, free_dofs, dirichlet_dofs,
                                        force_list, E0, Emin, p, nu0, rho, u_all,
                                        solver='auto', solver_config=None,
                                        elem_func: Callable = composer.simp_interpolation,
                                        rtol: float = 1e-5, maxiter=None,
                                        petsc_options=None, timer=None) -> np.ndarray:
    """Compliance f_i . u_i of every load case (reference :398-467)."""
    cfg = solver_config if solver_config is not None else \
        normalize_linear_solver_config(solver, rtol=rtol, maxiter=maxiter,
                                       petsc_options=petsc_options)
    comp, _ = _run_loads(basis, dirichlet_dofs, force_list, E0, Emin, p, nu0,
                         rho, u_all, elem_func, cfg, timer)
    return comp


def strain_energy_skfem_multi(basis, rho, U, E0, Emin, p, nu,
                              elem_func: Callable = composer.simp_interpolation):
    """Element strain energies U_e = 1/2 u_e^T K_e(rho) u_e, shape
    (n_elements, n_loads) (reference :507-537)."""
    dm = dev.device_mesh(basis.mesh)
    ke0 = dm.unit_ke(KE_ELASTIC, basis.X, basis.W, nu=nu)
    on_dev = _is_dev(U)
    E = dev.interpolate_modulus(dev.to_dev(rho), E0, Emin, p,
                                ramp=composer.is_ramp(elem_func))
    U2 = U if U.ndim == 2 else U[:, None]
    n_loads = U2.shape[1]
    out = torch.empty((n_loads, dm.n_elem), dtype=dev.F64, device="cuda")
    for i in range(n_loads):
        ui = U2[:, i].contiguous() if on_dev else dev.to_dev(np.ascontiguousarray(U2[:, i]))
        dm.element_energy(3, ke0, E, ui, out=out[i])
    res = out.t()
    return res if on_dev else res.cpu().numpy()


def strain_energy_skfem(basis, rho, u, E0, Emin, p, nu,
                        elem_func: Callable = composer.simp_interpolation):
    e = strain_energy_skfem_multi(basis, rho, u[:, None], E0, Emin, p, nu, elem_func)
    return e[:, 0]


class FEM_SimpLinearElasticity():
    """Linear-elastic FEM with SIMP / RAMP interpolated modulus (reference
    :540-676): ``objectives_multi_load`` assembles K(rho), enforces the Dirichlet
    set, solves every load case and returns the compliances;
    ``energy_multi_load`` returns the element strain energies."""

    def __init__(self, task, E_min_coeff: float,
                 density_interpolation: Callable = composer.simp_interpolation,
                 solver_config: LinearSolverConfig | None = None,
                 solver_option: Literal["spsolve", "cg_pyamg", "petsc", "petsc_spdirect"] = "spsolve",
                 petsc_options: PETScOptions | None = None):
        self.task = task
        self.E_max = task.E * 1.0
        self.E_min = task.E * E_min_coeff
        self.density_interpolation = density_interpolation
        self.solver_config = (
            solver_config if solver_config is not None else
            normalize_linear_solver_config(solver_option, petsc_options=petsc_options)
        )
        self.solver_option = self.solver_config.solver
        self.petsc_options = self.solver_config.petsc_options
        self._force_dev = None
        self._scaled = None

    def _forces(self, force_scale):
        fl = self.task.neumann_linear if isinstance(self.task.neumann_linear, list) \
            else [self.task.neumann_linear]
        if self._force_dev is None or len(self._force_dev) != len(fl):
            self._force_dev = [dev.to_dev(f) for f in fl]
            self._scaled = [torch.empty_like(f) for f in self._force_dev]
        if force_scale == 1.0:
            return self._force_dev
        for f, s in zip(self._force_dev, self._scaled):
            torch.mul(f, float(force_scale), out=s)
        return self._scaled

    @property
    def engine(self):
        return get_engine(self.task.basis, self.task.dirichlet_dofs, KE_ELASTIC,
                          self.task.nu)

    def objectives_multi_load(self, rho, p: float, u_dofs, timer=None,
                              force_scale: float = 1.0) -> np.ndarray:
        return compute_compliance_basis_multi_load(
            self.task.basis, self.task.free_dofs, self.task.dirichlet_dofs,
            self._forces(force_scale), self.E_max, self.E_min, p, self.task.nu,
            rho, u_dofs, elem_func=self.density_interpolation,
            solver_config=self.solver_config, timer=timer,
        )

    def energy_multi_load(self, rho, p: float, u_dofs):
        return strain_energy_skfem_multi(
            self.task.basis, rho, u_dofs, self.E_max, self.E_min, p, self.task.nu,
            elem_func=self.density_interpolation,
        )
