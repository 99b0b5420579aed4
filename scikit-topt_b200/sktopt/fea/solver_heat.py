"""Steady heat conduction with SIMP conductivity and Robin terms on the GPU.

Follows reference ``fea/solver_heat.py`` for ``objective="compliance"``
(SURVEY.md row a17):

* K = K_cond(rho) (:220, ``composer.assemble_conduction_matrix``)
      + sum of the task's real Robin facet matrices  int_G h u v  (:221-225)
      + the "virtual" Robin domain form  int_O h rho_n^p (1-rho_n)^q |grad rho_n| u v
        with rho_n the plain element->node average (:552-572, :723-745);
* emit = real Robin loads (h T_env) + virtual Robin load (:226-235, :565-571);
* per Dirichlet value set: enforce with T_D = value, solve (:160-192);
* J_i = T_i^T K T_i with the un-enforced K (:786-788); lambda = -2 T (:789);
* U_e = int_e 1/2 k_e |grad T|^2 (:256-303);
* dJ/drho = dC_drho_simp(rho, U, k_max, k_min, p) + explicit Robin term: nodal
  linear form h (a' |grad rho| phi v + a phi grad rho/max(|grad rho|,1e-12) . grad v),
  a = rho^p (1-rho)^q, phi = 2 T_env T - T^2, mapped node->element by
  sum_a g[t_a] / count[t_a] (:575-625, :928-980).

The reference solves with a sparse LU; here the enforced system goes through
the device Jacobi-PCG.

``objective="heat_exchange"`` (:306-324, :385-446, :791-869) and
``"averaged_temp"`` (:871-884) reuse the same operator: J = J_num / J_den on the
interface measure |grad rho_n| (``CellBasis`` with skfem's *default* quadrature,
not the task's), resp. J = sum(T); the adjoint systems K_e lambda = dJ/dT are
enforced with the *same* Dirichlet values as the state (reference quirk kept:
``solve_heat_system_multi`` is called with ``dirichlet_values_list`` :838-845),
and ``energy_multi_load`` returns the unit-conductivity elemental integrals of
grad T . grad lambda (:327-383, :518-549).  ``compliance_sensitivity_multi_load``
stays the SIMP conduction term for those objectives (:962-963), exactly like
the reference, whose optimiser always prefers it (common_density.py:1069-1082).
"""
from __future__ import annotations

from typing import Callable, Literal

import numpy as np
import torch

from sktopt._b200 import device as dev
from sktopt.fea import composer
from sktopt.fea._engine import KE_LAPLACE, get_engine
from sktopt.fea.solver_elastic import (
    LinearSolverConfig, normalize_linear_solver_config, _is_dev, _check_solver,
)
from sktopt.tools.logconf import mylogger

logger = mylogger(__name__)

HeatSolver = Literal["spsolve", "petsc", "petsc_spdirect"]
HeatSolverSelector = Literal["auto", "spsolve", "petsc", "petsc_spdirect"]


def _as_list(x):
    return x if isinstance(x, list) else [x]


# ------------------------------------------------ function-level API --------
def heat_energy_skfem_multi(basis, rho, T_all, k0, kmin, p,
                            elem_func: Callable = composer.simp_interpolation):
    """U_e = int_e 1/2 k_e |grad T|^2 for every temperature column, shape
    (n_elements, n_loads) (reference :256-303)."""
    dm = dev.device_mesh(basis.mesh)
    ke0 = dm.unit_ke(KE_LAPLACE, basis.X, basis.W)
    k = dev.interpolate_modulus(dev.to_dev(rho), k0, kmin, p, ramp=composer.is_ramp(elem_func))
    on_dev = _is_dev(T_all)
    T2 = T_all if T_all.ndim == 2 else T_all[:, None]
    out = torch.empty((T2.shape[1], dm.n_elem), dtype=dev.F64, device="cuda")
    for i in range(T2.shape[1]):
        Ti = T2[:, i].contiguous() if on_dev else dev.to_dev(np.ascontiguousarray(T2[:, i]))
        dm.element_energy(1, ke0, k, Ti, out=out[i])
    res = out.t()
    return res if on_dev else res.cpu().numpy()


def heat_energy_skfem(basis, rho, T, k0, kmin, p,
                      elem_func: Callable = composer.simp_interpolation):
    return heat_energy_skfem_multi(basis, rho, T[:, None], k0, kmin, p, elem_func)[:, 0]


def heat_exchange_grad_density_multi(basis, T_all, λ_all):
    """Elemental integrals of grad T . grad lambda, shape (n_elements, n_loads)
    (reference :343-383; ``avg_temp_grad_density_multi`` :536-549 is the same
    functional)."""
    dm = dev.device_mesh(basis.mesh)
    ke0 = dm.unit_ke(KE_LAPLACE, basis.X, basis.W)
    on_dev = _is_dev(T_all)
    T2 = T_all if T_all.ndim == 2 else T_all[:, None]
    L2 = λ_all if λ_all.ndim == 2 else λ_all[:, None]
    out = torch.empty((T2.shape[1], dm.n_elem), dtype=dev.F64, device="cuda")
    for i in range(T2.shape[1]):
        Ti = T2[:, i].contiguous() if on_dev else dev.to_dev(np.ascontiguousarray(T2[:, i]))
        Li = (L2[:, i].contiguous() if _is_dev(L2)
              else dev.to_dev(np.ascontiguousarray(L2[:, i])))
        dm.element_bilinear(1, ke0, None, Ti, Li, 1.0, out=out[i])
    res = out.t()
    return res if on_dev else res.cpu().numpy()


avg_temp_grad_density_multi = heat_exchange_grad_density_multi


class _HeatDevice:
    """Device buffers of one heat task (scalar CSR on the node graph)."""

    def __init__(self, task):
        self.task = task
        basis = task.basis
        d_nodes = _as_list(task.dirichlet_nodes)
        # One enforced operator serves every load case, so the load cases must fix
        # the same nodes (their VALUES may differ).  The reference enforces per
        # load with D_i (fea/solver_heat.py:160-175); different node sets would
        # need one operator per load and are rejected instead of being solved with
        # the wrong boundary rows.
        d0 = np.unique(np.asarray(d_nodes[0], dtype=np.int64))
        for dn in d_nodes[1:]:
            if not np.array_equal(d0, np.unique(np.asarray(dn, dtype=np.int64))):
                raise NotImplementedError(
                    "heat load cases with different Dirichlet node sets are not supported: "
                    "all load cases must fix the same nodes (values may differ)")
        # the heat operator is replicated on every rank (scalar system, small)
        self.eng = get_engine(basis, np.asarray(d_nodes[0], dtype=np.int64), KE_LAPLACE,
                              shard=False)
        eng = self.eng
        dm = eng.dm
        n = eng.n_dof
        self.K = eng.vals                                  # un-enforced total matrix
        self.K_e = torch.empty_like(self.K)                # enforced copy for the solve
        self.V = torch.empty_like(self.K)                  # virtual Robin matrix
        self.ones = torch.ones(n, dtype=dev.F64, device="cuda")
        self.tmp = torch.empty(n, dtype=dev.F64, device="cuda")
        self.emit = torch.empty(n, dtype=dev.F64, device="cuda")
        self.rho_n = torch.empty(n, dtype=dev.F64, device="cuda")
        self.count = dm.e2n_wsum(None)                     # elements per node (0 -> 1)
        self.xD = []
        for nodes, val in zip(d_nodes, _as_list(task.dirichlet_values)):
            x = np.zeros(n)
            x[np.asarray(nodes, dtype=np.int64)] = float(val)
            self.xD.append(dev.to_dev(x))
        # real Robin facet terms: rho-independent in the reference
        rp, ci = dm.node_graph()
        keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp)) * n + ci
        base = np.zeros(keys.size)
        for B in (task.robin_bilinear or []):
            B = B.tocsr()
            B.sort_indices()
            bk = np.repeat(np.arange(n, dtype=np.int64), np.diff(B.indptr)) * n + B.indices
            pos = np.searchsorted(keys, bk)
            if not np.array_equal(keys[pos], bk):
                raise ValueError("Robin matrix couples nodes outside the mesh graph")
            base[pos] += B.data
        self.base = dev.to_dev(base) if (task.robin_bilinear or []) else None
        emit0 = np.zeros(n)
        for f in (task.robin_linear or []):
            emit0 = emit0 + f
        self.emit_real = dev.to_dev(emit0)
        self.tables = dm.geom_tables(basis.X, basis.W)
        self.scale_v = None
        self._cell_tables = None
        self.ones_e = None

    def cell_tables(self):
        """Quadrature tables of ``skfem.CellBasis(mesh, elem)`` with the default
        integration order (reference :793), used by the heat_exchange forms."""
        if self._cell_tables is None:
            from sktopt._fem.basis import Basis
            cb = Basis(self.task.basis.mesh, self.task.basis.elem)
            self._cell_tables = self.eng.dm.geom_tables(cb.X, cb.W)
        return self._cell_tables


class FEM_SimpLinearHeatConduction():
    """Heat-conduction FEM façade (reference :628-980)."""

    def __init__(self, task, E_min_coeff: float,
                 density_interpolation: Callable = composer.simp_interpolation,
                 solver_config: LinearSolverConfig | None = None,
                 solver_option: Literal["spsolve", "petsc", "petsc_spdirect"] = "spsolve",
                 q: int = 4):
        self.task = task
        self.k_max = task.k * 1.0
        self.k_min = task.k * E_min_coeff
        self.density_interpolation = density_interpolation
        self.solver_config = (
            solver_config if solver_config is not None else
            normalize_linear_solver_config(solver_option)
        )
        self.solver_option = self.solver_config.solver
        self.λ_all = None
        self.q = q
        self._warned_robin_compliance = False
        self._dev = None

    def _device(self) -> _HeatDevice:
        if self._dev is None:
            self._dev = _HeatDevice(self.task)
        return self._dev

    @property
    def engine(self):
        return self._device().eng

    def _check_objective(self):
        if self.task.objective not in ("compliance", "heat_exchange", "averaged_temp"):
            raise ValueError(f"Unknown objective: {self.task.objective}")

    def _robin_scalar(self):
        h, T_env = self.task.robin_coefficient, self.task.robin_bc_value
        if isinstance(h, list) or isinstance(T_env, list):
            # the reference multiplies by h directly, so lists break there too
            raise TypeError("virtual Robin terms need scalar robin_coefficient / robin_bc_value")
        return float(h), float(T_env)

    def objectives_multi_load(self, rho, p: float, u_dofs, timer=None,
                              force_scale: float = 1.0) -> np.ndarray:
        self._check_objective()
        _check_solver(self.solver_config.solver)
        st = self._device()
        eng, dm = st.eng, st.eng.dm
        rho_d = dev.to_dev(rho)
        ramp = composer.is_ramp(self.density_interpolation)
        eng.set_modulus(rho_d, self.k_max, self.k_min, p, ramp=ramp)
        eng.assemble(enforce=False)                          # K_cond -> st.K
        st.emit.copy_(st.emit_real)
        if st.base is not None:
            dev.axpby(1.0, st.base, 1.0, st.K)
        if self.task.robin_coefficient is not None:
            h, T_env = self._robin_scalar()
            dm.e2n(None, rho_d, None, 0.0, st.count, out=st.rho_n)
            N, G, dx, Mq = st.tables
            st.scale_v = dm.robin_virtual_scale(st.tables, st.rho_n, h, p, self.q,
                                                out=st.scale_v)
            dm.assemble_terms(Mq, st.scale_v, out=st.V)
            dev.axpby(1.0, st.V, 1.0, st.K)
            # virtual load = T_env * V * 1 (shape functions sum to one)
            dev.spmv(eng.row_ptr, eng.col_idx, st.V, st.ones, 1, out=st.tmp)
            dev.axpby(T_env, st.tmp, 1.0, st.emit)
            if self.task.objective == "compliance" and not self._warned_robin_compliance:
                logger.warning(
                    "Heat objective='compliance' with Robin boundaries evaluates "
                    "T^T K T, which includes Dirichlet reaction work.")
                self._warned_robin_compliance = True
        st.K_e.copy_(st.K)
        if eng.has_dirichlet:
            dev.csr_enforce(eng.row_ptr, eng.col_idx, st.K_e, eng.dir_mask)
        eng.update_preconditioner(st.K_e)

        n_loads = len(st.xD)
        Ts = []
        for i in range(n_loads):
            T = self._solve_enforced(st.emit, i, slot=i)
            Ts.append(T)
            self._store(u_dofs, i, T)

        objective = self.task.objective
        if objective == "compliance":
            J = np.empty(n_loads)
            for i, T in enumerate(Ts):
                dev.spmv(eng.row_ptr, eng.col_idx, st.K, T, 1, out=st.tmp)
                J[i] = dev.dot(T, st.tmp)
            self.λ_all = -2.0 * u_dofs
            return J
        if objective == "heat_exchange":
            return self._heat_exchange(st, rho_d, p, u_dofs, Ts)
        # averaged_temp (:871-884): J = sum(T), adjoint load = ones
        J = np.array([dev.dot(T, st.ones) for T in Ts])
        lam = self._new_like(u_dofs)
        for i in range(n_loads):
            self._store(lam, i, self._solve_enforced(st.ones, i, slot=n_loads + i))
        self.λ_all = lam
        return J

    # -- helpers of objectives_multi_load --------------------------------------
    def _solve_enforced(self, load_vec, i: int, slot: int):
        """K_e x = enforce(load_vec; D, x_D = Dirichlet value of load i)
        (``solve_heat_system_multi`` :136-200); ``slot`` selects the warm-start
        vector (state and adjoint solves keep separate ones)."""
        st = self._device()
        eng = st.eng
        dev.spmv(eng.row_ptr, eng.col_idx, st.K, st.xD[i], 1, out=st.tmp)
        dev.enforce_rhs(load_vec, st.tmp, eng.dir_mask, st.xD[i], out=eng.rhs)
        # start from the exact Dirichlet values: their identity rows then
        # have zero residual for the whole solve (as exact as the LU)
        x0 = eng.solution(slot)
        dev.enforce_rhs(x0, None, eng.dir_mask, st.xD[i], out=x0)
        return eng.solve(eng.rhs, slot, self.solver_config.rtol,
                         self.solver_config.maxiter, vals=st.K_e)

    @staticmethod
    def _store(dst, i: int, vec):
        if _is_dev(dst):
            (dst[:, i] if dst.ndim == 2 else dst).copy_(vec)
        else:
            dst[:, i] = vec.cpu().numpy()

    @staticmethod
    def _new_like(u_dofs):
        return torch.zeros_like(u_dofs) if _is_dev(u_dofs) else np.zeros_like(u_dofs)

    def _heat_exchange(self, st, rho_d, p, u_dofs, Ts):
        """heat_exchange branch of objectives_multi_load (:791-869)."""
        if self.task.robin_coefficient is None:
            raise RuntimeError("heat_exchange objective requires Robin boundary data.")
        eng, dm = st.eng, st.eng.dm
        h, T_env = self._robin_scalar()
        n_loads = len(Ts)
        tab = st.cell_tables()
        # st.rho_n was filled by the virtual Robin assembly above; the kernel
        # forms h_eff nodal = h rho_n^p (1 - rho_n)^q itself (:825-826)
        den_e, _, local = dm.heat_exchange_local(tab, st.rho_n, None, p, self.q, h, T_env,
                                                 want_num=False, want_local=True)
        if st.ones_e is None:
            st.ones_e = torch.ones_like(den_e)
        ones_e = st.ones_e
        J_den = dev.dot(den_e, ones_e)
        if J_den <= 1e-16:
            self.λ_all = self._new_like(u_dofs)
            return np.zeros(n_loads)
        J = np.empty(n_loads)
        for i, T in enumerate(Ts):
            _, num_e, _ = dm.heat_exchange_local(tab, st.rho_n, T, p, self.q, h, T_env,
                                                 want_num=True, want_local=False)
            J[i] = dev.dot(num_e, ones_e) / J_den
        rhs = torch.empty_like(st.tmp)
        dev.affine(1.0 / J_den, dm.local_to_nodes(local), 0.0, None, 0.0, rhs)   # dJ/dT (:830)
        lam = self._new_like(u_dofs)
        lam_d = [self._solve_enforced(rhs, i, slot=n_loads + i) for i in range(n_loads)]
        w = float(getattr(self.task, "avg_temp_weight", 0.0))
        if w != 0.0:
            avg = np.array([dev.dot(T, st.ones) for T in Ts])
            lam_avg = [self._solve_enforced(st.ones, i, slot=2 * n_loads + i)
                       for i in range(n_loads)]
            if n_loads > 1:
                hx_scale = max(float(np.mean(np.abs(J))), 1.0e-12)
                avg_scale = max(float(np.mean(np.abs(avg))), 1.0e-12)
                J = J / hx_scale + w * (avg / avg_scale)
                mix = (1.0 / hx_scale, w / avg_scale)
            else:
                J = J + w * avg
                mix = (1.0, w)
            lam_d = [dev.affine(mix[0], l, mix[1], la, 0.0, torch.empty_like(l))
                     for l, la in zip(lam_d, lam_avg)]
        for i, l in enumerate(lam_d):
            self._store(lam, i, l)
        self.λ_all = lam
        return J

    def _energy_dev(self, rho_d, p, u_dofs):
        st = self._device()
        eng = st.eng
        on_dev = _is_dev(u_dofs)
        eng.set_modulus(rho_d, self.k_max, self.k_min, p,
                        ramp=composer.is_ramp(self.density_interpolation))
        U2 = u_dofs if u_dofs.ndim == 2 else u_dofs[:, None]
        out = torch.empty((U2.shape[1], eng.n_elem), dtype=dev.F64, device="cuda")
        Ts = []
        for i in range(U2.shape[1]):
            Ti = U2[:, i].contiguous() if on_dev else dev.to_dev(np.ascontiguousarray(U2[:, i]))
            eng.energy(Ti, out=out[i])
            Ts.append(Ti)
        return out, Ts

    def energy_multi_load(self, rho, p: float, u_dofs):
        self._check_objective()
        if self.task.objective == "compliance":
            out, _ = self._energy_dev(dev.to_dev(rho), p, u_dofs)
            return out.t() if _is_dev(u_dofs) else out.t().cpu().numpy()
        # heat_exchange / averaged_temp: elemental integrals of grad T . grad lambda
        # with unit conductivity (:327-383, :518-549, :886-917)
        if self.λ_all is None:
            raise RuntimeError("adjoint field λ_all is not computed.")
        eng = self._device().eng
        on_dev = _is_dev(u_dofs)
        U2 = u_dofs if u_dofs.ndim == 2 else u_dofs[:, None]
        L2 = self.λ_all if self.λ_all.ndim == 2 else self.λ_all[:, None]
        out = torch.empty((U2.shape[1], eng.n_elem), dtype=dev.F64, device="cuda")
        for i in range(U2.shape[1]):
            Ti = U2[:, i].contiguous() if on_dev else dev.to_dev(np.ascontiguousarray(U2[:, i]))
            Li = (L2[:, i].contiguous() if _is_dev(L2)
                  else dev.to_dev(np.ascontiguousarray(L2[:, i])))
            eng.dm.element_bilinear(1, eng.unit_ke, None, Ti, Li, 1.0, out=out[i])
        return out.t() if on_dev else out.t().cpu().numpy()

    def compliance_sensitivity_multi_load(self, rho, p: float, u_dofs):
        st = self._device()
        eng, dm = st.eng, st.eng.dm
        rho_d = dev.to_dev(rho)
        ramp = composer.is_ramp(self.density_interpolation)
        energy, Ts = self._energy_dev(rho_d, p, u_dofs)
        grad = torch.empty_like(energy)
        for i in range(energy.shape[0]):
            dev.dc_drho(rho_d, energy[i], self.k_max, self.k_min, p, ramp=ramp,
                        out=grad[i])
        if self.task.objective == "compliance" and self.task.robin_coefficient is not None:
            h, T_env = self._robin_scalar()
            dm.e2n(None, rho_d, None, 0.0, st.count, out=st.rho_n)
            local = nodal = elem = None
            for i, Ti in enumerate(Ts):
                local = dm.robin_explicit_local(st.tables, st.rho_n, Ti, h, T_env,
                                                p, self.q, out=local)
                nodal = dm.local_to_nodes(local, divisor=st.count, out=nodal)
                elem = dm.n2e_mean(nodal, out=elem)
                dev.axpby(float(dm.nen), elem, 1.0, grad[i])    # sum over the element's nodes
        res = grad.t()
        return res if _is_dev(u_dofs) else res.cpu().numpy()
