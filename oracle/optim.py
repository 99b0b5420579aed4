"""CPU ORACLE (test infrastructure): projection, sensitivities, OC / LogMOC
updates and the optimiser loop, NumPy only.  PARITY UNPINNED (see fem.py).

Restates core/projection.py:27-118, core/derivatives.py:25-68,
core/optimizers/oc.py:16-95,149-250, core/optimizers/logmoc.py:36-66,90-258,
core/optimizers/common_density.py:711-745,1014-1134 and the schedule
functions tools/scheduler.py:74-261,809-830."""
from __future__ import annotations

import time

import numpy as np

from . import fem
from .filters import HelmholtzOracle, SpatialOracle


# ------------------------------------------------------------ schedules ---
def sched_step(it, total, init, target, num_steps, curvature=None, mode="linear"):
    """schedule_step / _accelerating / _decelerating (scheduler.py:74-216)
    behind Scheduler.value's guards (:809-830)."""
    if num_steps is None or num_steps < 0 or it >= total:
        return target
    if num_steps <= 1:
        return target
    idx = min(int(it // (total / num_steps)), num_steps - 1)
    a = idx / (num_steps - 1)
    if mode == "accelerating":
        a = a ** curvature
    elif mode == "decelerating":
        a = 1 - (1 - a) ** curvature
    return (1 - a) * init + a * target


def sched_sawtooth(it, total, init, target, num_steps):
    """schedule_sawtooth_decay (scheduler.py:219-261) behind Scheduler.value."""
    if it >= total:
        return target
    it0 = it - 1
    size = total / num_steps
    local = it0 - int(it0 // size) * size
    a = min(local / size, 1.0)
    return (1 - a) * init + a * target


# ----------------------------------------------------------- projection ---
def heaviside(rho, beta, eta):
    den = np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta)) + 1e-12
    return (np.tanh(beta * eta) + np.tanh(beta * (rho - eta))) / den


def heaviside_derivative(rho, beta, eta):
    den = np.tanh(beta * eta) + np.tanh(beta * (1.0 - eta)) + 1e-12
    return beta / np.cosh(beta * (rho - eta)) ** 2 / den


def dC_drho_simp(rho, U, E0, Emin, p):
    """core/derivatives.py:25-52 (clamps rho >= 1e-6, E >= 1e-12)."""
    rc = np.maximum(rho, 1e-6)
    dE = p * (E0 - Emin) * rc ** (p - 1)
    E = Emin + (E0 - Emin) * rc ** p
    return -2.0 * U * dE / np.maximum(E, 1e-12)


# ------------------------------------------------------------------- OC ---
def dC_drho_ramp(rho, U, E0, Emin, p):
    """core/derivatives.py:56-68 (RAMP: no clamping of rho)."""
    den = 1.0 + p * (1.0 - rho)
    dE = (E0 - Emin) * (den - p * rho) / den ** 2
    E = Emin + (E0 - Emin) * (rho / den)
    return -2.0 * U * dE / np.maximum(E, 1e-12)


def oc_bisection(dC, rho_e, rho_full, design, filt, rho_min, rho_max, move, eta,
                 eps, vol_frac, beta, beta_eta, vol_d, vol_sum, sr_min, sr_max,
                 max_iter=1000, tolerance=1e-5, vol_tol=1e-4, l1=1e-7, l2=1e7):
    """bisection_with_physical_volume (core/optimizers/oc.py:16-95)."""
    it = 0
    lmid = 0.5 * (l1 + l2)
    steps = 0
    while True:
        sr = np.clip((-dC / (lmid + eps)) ** eta, sr_min, sr_max)
        cand = np.clip(rho_e * sr, np.maximum(rho_e - move, rho_min),
                       np.minimum(rho_e + move, rho_max))
        full = rho_full.copy()
        full[design] = cand
        proj = heaviside(filt.forward(full), beta, beta_eta)
        vol_error = np.sum(proj[design] * vol_d) / vol_sum - vol_frac
        steps += 1
        if abs(vol_error) < vol_tol or it >= max_iter or abs(l2 - l1) <= tolerance:
            break
        if vol_error > 0:
            l1 = lmid
        else:
            l2 = lmid
        it += 1
        lmid = 0.5 * (l1 + l2)
    return cand, sr, lmid, vol_error, steps


# --------------------------------------------------------------- LogMOC ---
def logmoc_step(rho, dL, eta, move, rho_min, rho_max, clip):
    """lagrangian_log_update (core/optimizers/logmoc.py:36-66)."""
    g = np.clip(dL, -clip, clip)
    r = np.clip(rho, rho_min, rho_max)
    lr = np.log(r)
    w = np.log(move / np.exp(lr) + 1.0)
    lo = lr - w
    hi = lo + 2.0 * w
    return np.clip(np.exp(np.clip(lr - eta * g, lo, hi)), rho_min, rho_max)


# ----------------------------------------------------------------- loop ---
class Problem:
    """Arrays of one elasticity task (see oracle.mesh.toy_base)."""

    def __init__(self, p, t, dirichlet_dofs, forces, design, pinned, volumes,
                 E, nu, fixed=None, intorder=2):
        self.p, self.t = p, t
        self.D = np.asarray(dirichlet_dofs)
        self.forces = forces if isinstance(forces, list) else [forces]
        self.design = np.asarray(design)
        self.pinned = np.asarray(pinned)
        self.fixed = np.asarray(fixed) if fixed is not None else np.array([], dtype=int)
        self.vol = volumes
        self.E, self.nu, self.intorder = E, nu, intorder
        self.design_mask = np.isin(np.arange(t.shape[1]), self.design)


def run(problem: Problem, method="oc", max_iters=5, filter_type="helmholtz",
        filter_radius=0.01, vol_frac=0.8, p_sched=(1.0, 3.0, 3),
        beta_sched=(1.0, 2.0, 3, 2.0), move_sched=(0.3, 0.1, 6), eta=None,
        rho_min=1e-2, rho_max=1.0, E_min_coeff=1e-3, beta_eta=0.5,
        solver="spsolve", rtol=1e-8, cg_maxiter=None, lambda_lower=1e-7,
        lambda_upper=1e7, logmoc=None, iters=None, timings=None, step_times=None,
        interpolation="SIMP", sensitivity_filter=False, backend=None,
        filter_solver="splu", time_budget=None, filter_matrices=None):
    """DensityMethod._optimize_impl (common_density.py:1014-1134) with the
    default schedules of DensityMethodConfig / OC_Config / LogMOC_Config.
    Returns dict(rho, compliance[], vol_error[], rho_hist[]).

    ``backend`` (``oracle.cport.CBackend``): assembly + enforce, the Jacobi-PCG
    and the element energies run through the C / OpenMP port (same algorithms on
    all host cores: the CPU baseline at BASELINE's full sizes); ``filter_solver=
    'cg'`` solves the Helmholtz systems by cg instead of a sparse LU.
    ``time_budget`` (seconds): stop after the first iteration that ends later
    than this (timed runs on hosts of unknown speed; at least one iteration)."""
    pr = problem
    ne = pr.t.shape[1]
    E0, Emin = pr.E, pr.E * E_min_coeff
    # interpolation_funcs (common_density.py:392-404)
    interp, dC_drho = {"SIMP": (fem.simp, dC_drho_simp),
                       "RAMP": (fem.ramp, dC_drho_ramp)}[interpolation]
    if filter_type == "helmholtz":
        filt = HelmholtzOracle(pr.p, pr.t, pr.vol, pr.design_mask, solver=filter_solver,
                               matrices=filter_matrices)
    else:
        filt = SpatialOracle(pr.p, pr.t, pr.design_mask)
    filt.set_radius(filter_radius)
    if eta is None:
        eta = 0.5 if method == "oc" else 0.6
    lm = dict(mu_p=5.0, lambda_v=0.1, lambda_decay=0.90, lambda_lower=-1e7,
              lambda_upper=1e7, lagrangian_clip=1.0, lagrangian_percentile=95.0,
              lagrangian_scale_floor=1e-8)
    if logmoc:
        lm.update(logmoc)
    # initial density (common_density.py:711-745)
    rho = np.clip(np.zeros(ne) + vol_frac, rho_min, rho_max)
    rho[pr.pinned] = 1.0
    rho[pr.fixed] = 1.0
    vol_d = pr.vol[pr.design]
    vol_sum = float(np.sum(vol_d))
    dV_design = vol_d / vol_sum
    running_scale = 0.0
    lambda_v = lm["lambda_v"]
    hist = dict(compliance=[], vol_error=[], rho=[], cg_iters=[], bisection_steps=[])
    tm = timings if timings is not None else {}

    def tick(name, t0):
        tm[name] = tm.get(name, 0.0) + time.perf_counter() - t0

    n_it = max_iters if iters is None else iters
    t_start = time.perf_counter()
    for it in range(1, n_it + 1):
        if step_times is not None:
            step_times.append(time.perf_counter())      # start of every iteration
        pw = sched_step(it, max_iters, *p_sched)
        beta = sched_step(it, max_iters, beta_sched[0], beta_sched[1], beta_sched[2],
                          beta_sched[3], "accelerating")
        move = sched_sawtooth(it, max_iters, *move_sched)
        t0 = time.perf_counter()
        rho_f = filt.forward(rho)
        rho_p = heaviside(rho_f, beta, beta_eta)
        tick("filter_and_project", t0)
        t0 = time.perf_counter()
        if backend is not None:
            # homogeneous Dirichlet data: enforce is folded into the assembly
            K_e = backend.assemble(interp(rho_p, E0, Emin, pw), enforce=True)
        else:
            K = fem.assemble_stiffness(pr.p, pr.t, rho_p, E0, Emin, pw, pr.nu, pr.intorder,
                                       interp=interp)
        tick("assemble", t0)
        t0 = time.perf_counter()
        if backend is None:
            K_e, _ = fem.enforce(K, pr.forces[0], pr.D)
        tick("enforce_bc", t0)
        t0 = time.perf_counter()
        comps, U = [], []
        for f in pr.forces:
            F_e = np.array(f, dtype=float)
            F_e[pr.D] = 0.0
            if backend is not None:
                u, nit, _ = backend.pcg(F_e, rtol, cg_maxiter)
            else:
                u, _, nit = fem.solve(K_e, F_e, solver, rtol, cg_maxiter)
            hist["cg_iters"].append(nit)
            comps.append(float(F_e @ u))
            U.append(u)
        U = np.column_stack(U)
        tick("solve", t0)
        compliance = float(np.mean(comps))
        t0 = time.perf_counter()
        if backend is not None:
            sc = interp(rho_p, E0, Emin, pw)
            energy = np.column_stack([backend.energy(sc, U[:, l]) for l in range(U.shape[1])])
        else:
            energy = fem.strain_energy(pr.p, pr.t, rho_p, U, E0, Emin, pw, pr.nu, pr.intorder,
                                       interp=interp)
        tick("energy", t0)
        t0 = time.perf_counter()
        dC_full = np.zeros(ne)
        dH = heaviside_derivative(rho_f, beta, beta_eta)
        for l in range(U.shape[1]):
            dC_full += filt.gradient(dC_drho(rho_p, energy[:, l], E0, Emin, pw) * dH)
        dC_full /= U.shape[1]
        if sensitivity_filter:                            # common_density.py:1102-1106
            dC_full = filt.forward(dC_full)
        tick("sensitivity", t0)
        dC = dC_full[pr.design].copy()
        rho_e = rho[pr.design].copy()
        t0 = time.perf_counter()
        if method == "oc":
            scale = max(np.max(np.abs(dC)), 1e-12)
            running_scale = 0.6 * running_scale + 0.4 * scale if it > 1 else scale
            dC = dC / running_scale
            new, _, lmid, vol_error, steps = oc_bisection(
                dC, rho_e, rho, pr.design, filt, rho_min, rho_max, move, eta,
                1e-12, vol_frac, beta, beta_eta, vol_d, vol_sum, 0.7, 1.3,
                l1=lambda_lower, l2=lambda_upper)
            hist["bisection_steps"].append(steps)
        else:
            dVf = np.zeros(ne)
            dVf[pr.design] = dV_design
            dVf = heaviside_derivative(rho_f, beta, beta_eta) * dVf
            back = filt.gradient(dVf)
            if np.allclose(back, 0.0) and np.any(dVf > 0.0):
                back = filt.forward(dVf)
            dV_chain = back[pr.design]
            vol_error = np.sum(rho_p[pr.design] * vol_d) / vol_sum - vol_frac
            penalty = lm["mu_p"] * vol_error
            lambda_v = (lm["lambda_decay"] * lambda_v + (1.0 - lm["lambda_decay"]) * penalty
                        if it > 1 else penalty)
            lambda_v = float(np.clip(lambda_v, lm["lambda_lower"], lm["lambda_upper"]))
            dL = dC + lambda_v * dV_chain
            scale = max(np.percentile(np.abs(dL), lm["lagrangian_percentile"]),
                        lm["lagrangian_scale_floor"])
            running_scale = 0.2 * running_scale + 0.8 * scale if it > 1 else scale
            dL = dL / running_scale
            new = logmoc_step(rho_e, dL, eta, move, rho_min, rho_max, lm["lagrangian_clip"])
        tick("rho_update", t0)
        rho[pr.design] = new
        rho[pr.pinned] = 1.0
        hist["compliance"].append(compliance)
        hist["vol_error"].append(float(vol_error))
        hist["rho"].append(rho[pr.design].copy())
        if time_budget is not None and time.perf_counter() - t_start > time_budget:
            break
    if step_times is not None:
        step_times.append(time.perf_counter())          # end of the last one
    hist["rho_final"] = rho
    hist["rho_projected"] = rho_p
    hist["energy"] = energy
    hist["u"] = U
    return hist
