"""CPU ORACLE (test infrastructure): heat-conduction path with Robin terms,
objectives "compliance", "heat_exchange" and "averaged_temp".  PARITY UNPINNED
(see fem.py).

Restates fea/solver_heat.py:136-253 (enforce + LU), :256-303 (energy),
:552-625 (virtual Robin forms, explicit sensitivity form, node<->element maps),
:705-789 (objectives_multi_load, compliance branch), :928-980
(compliance_sensitivity_multi_load) and mesh/task_heat.py:60-139 (real Robin
facet terms) with NumPy / SciPy; ``objectives`` adds :306-446 (heat-exchange
functionals and adjoint load), :791-884 (heat_exchange / averaged_temp branches)
and :327-383, :518-549 (grad T . grad lambda elemental integrals)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .optim import dC_drho_simp


def nodal_average(t, rho, n_nodes):
    """_element_to_nodal_average (fea/solver_heat.py:600-613): plain mean."""
    s = np.zeros(n_nodes)
    c = np.zeros(n_nodes)
    np.add.at(s, t.ravel(), np.repeat(rho[None, :], t.shape[0], axis=0).ravel())
    np.add.at(c, t.ravel(), 1.0)
    return s / np.maximum(c, 1.0), c


def _fields(p, t, rho_n, intorder):
    nen = t.shape[0]
    X, W = fem.quadrature(nen, intorder)
    N, G, dJ = fem.physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    re = rho_n[t.astype(np.int64)]                       # (nen, ne)
    rq = np.einsum("aq,ae->eq", N, re)
    gq = np.einsum("eqad,ae->eqd", G, re)
    return N, G, dx, rq, gq


def virtual_robin(p, t, rho_n, h, T_env, pw, q, intorder):
    """get_robin_virtual (fea/solver_heat.py:552-572): matrix and load."""
    N, G, dx, rq, gq = _fields(p, t, rho_n, intorder)
    iface = np.sqrt(np.sum(gq ** 2, axis=2))
    w = h * rq ** pw * (1.0 - rq) ** q * iface * dx      # (ne, nq)
    Ve = np.einsum("eq,aq,bq->eab", w, N, N)
    fe = T_env * np.einsum("eq,aq->ea", w, N)
    V = fem.assemble(t, Ve, 1, pattern=fem.structural_pattern(t, p.shape[1], 1))
    f = np.zeros(p.shape[1])
    np.add.at(f, t.T.astype(np.int64), fe)
    return V, f


def solve_compliance(p, t, rho, k0, kmin, pw, q, h, T_env, robin_B, robin_f,
                     D_nodes, D_value, intorder=2):
    """objectives_multi_load, compliance branch (fea/solver_heat.py:705-789),
    one Dirichlet value set.  Returns (J, T, K_total)."""
    n = p.shape[1]
    k = fem.simp(rho, k0, kmin, pw)
    K = fem.assemble_scalar(p, t, k, intorder, "laplace")
    emit = np.zeros(n)
    for B in robin_B:
        K = K + B
    for f in robin_f:
        emit = emit + f
    if h is not None:
        rho_n, _ = nodal_average(t, rho, n)
        V, fv = virtual_robin(p, t, rho_n, h, T_env, pw, q, intorder)
        K = K + V
        emit = emit + fv
    K = K.tocsr()
    xD = np.full(len(D_nodes), float(D_value))
    K_e, f_e = fem.enforce(K, emit, np.asarray(D_nodes), xD)
    T = spla.splu(K_e.tocsc()).solve(f_e)
    return float(T @ (K @ T)), T, K


def grad_dot_energy(p, t, T, lam, intorder=2):
    """heat_exchange_grad_density_multi / avg_temp_grad_density_multi
    (fea/solver_heat.py:327-383, :518-549): int_e grad T . grad lambda."""
    X, W = fem.quadrature(t.shape[0], intorder)
    N, G, dJ = fem.physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    ti = t.astype(np.int64)
    gT = np.einsum("eqad,ae->eqd", G, T[ti])
    gL = np.einsum("eqad,ae->eqd", G, lam[ti])
    return np.einsum("eqd,eqd,eq->e", gT, gL, dx)


def heat_exchange_forms(p, t, rho_n, T, h, T_env, pw, q):
    """(J_num, J_den, adjoint load before the division by J_den) on
    ``CellBasis(mesh, elem)`` with skfem's default integration order
    (fea/solver_heat.py:306-324, :385-446, :793-830)."""
    n = p.shape[1]
    N, G, dx, rq, gq = _fields(p, t, rho_n, fem.default_intorder(t.shape[0]))
    iface = np.sqrt(np.sum(gq ** 2, axis=2))
    J_den = float(np.sum(iface * dx))                     # :385-392, :796-797
    ti = t.astype(np.int64)
    Tq = np.einsum("aq,ae->eq", N, T[ti])
    J_num = float(np.sum(-T_env * h * (Tq - T_env) * iface * dx))     # :306-324
    heff_n = h * rho_n ** pw * (1.0 - rho_n) ** q          # :825-826 (nodal, then interpolated)
    hq = np.einsum("aq,ae->eq", N, heff_n[ti])
    fe = np.einsum("eq,aq->ea", -T_env * hq * iface * dx, N)          # :395-413
    rhs = np.zeros(n)
    np.add.at(rhs, ti.T, fe)
    return J_num, J_den, rhs


def objectives(p, t, rho, k0, kmin, pw, q, h, T_env, robin_B, robin_f,
               D_nodes, D_value, objective, intorder=2, avg_temp_weight=0.0):
    """objectives_multi_load (fea/solver_heat.py:705-889) for one Dirichlet value
    set and any of the three objectives.  Returns (J, T, lambda, K_total)."""
    n = p.shape[1]
    J, T, K = solve_compliance(p, t, rho, k0, kmin, pw, q, h, T_env, robin_B, robin_f,
                               D_nodes, D_value, intorder)
    if objective == "compliance":
        return J, T, -2.0 * T, K
    D = np.asarray(D_nodes)
    xD = np.full(len(D), float(D_value))
    emit = np.zeros(n)
    for f in robin_f:
        emit = emit + f

    def adjoint(rhs):
        # solve_heat_system_multi(K_csr, rhs, D, dirichlet_VALUES, ...) :838-845
        K_e, f_e = fem.enforce(K, rhs, D, xD)
        return spla.splu(K_e.tocsc()).solve(f_e)

    if objective == "averaged_temp":                      # :871-884
        return float(np.sum(T)), T, adjoint(np.ones(n)), K
    assert objective == "heat_exchange" and h is not None
    rho_n, _ = nodal_average(t, rho, n)
    J_num, J_den, rhs = heat_exchange_forms(p, t, rho_n, T, h, T_env, pw, q)
    if J_den <= 1e-16:
        return 0.0, T, np.zeros(n), K
    J = J_num / J_den
    rhs = rhs / J_den                                     # :830
    lam = adjoint(rhs)
    if avg_temp_weight != 0.0:                            # :847-869, single load
        J = J + avg_temp_weight * float(np.sum(T))
        lam = lam + avg_temp_weight * adjoint(np.ones(n))
    return J, T, lam, K


def sensitivity(p, t, rho, T, k0, kmin, pw, q, h, T_env, intorder=2):
    """compliance_sensitivity_multi_load (fea/solver_heat.py:928-980)."""
    U = fem.heat_energy(p, t, rho, T, k0, kmin, pw, intorder)[:, 0]
    g = dC_drho_simp(rho, U, k0, kmin, pw)
    if h is None:
        return g, U
    n = p.shape[1]
    rho_n, count = nodal_average(t, rho, n)
    N, G, dx, rq, gq = _fields(p, t, rho_n, intorder)
    Tq = np.einsum("aq,ae->eq", N, T[t.astype(np.int64)])
    iface = np.sqrt(np.sum(gq ** 2, axis=2))
    safe = np.maximum(iface, 1e-12)
    a = rq ** pw * (1.0 - rq) ** q
    da = pw * rq ** (pw - 1.0) * (1.0 - rq) ** q - q * rq ** pw * (1.0 - rq) ** (q - 1.0)
    phi = 2.0 * T_env * Tq - Tq * Tq
    term0 = np.einsum("eq,aq->ea", da * iface * phi * dx, N)
    term1 = np.einsum("eq,eqd,eqad->ea", a * phi / safe * dx, gq, G)
    nodal = np.zeros(n)
    np.add.at(nodal, t.T.astype(np.int64), h * (term0 + term1))
    elem = np.zeros(t.shape[1])
    for a_loc in range(t.shape[0]):                     # _nodal_gradient_to_element_gradient
        elem += nodal[t[a_loc]] / np.maximum(count[t[a_loc]], 1.0)
    return g + elem, U


def quad_facet_mass(p, cyc, coeff):
    """asm(BilinearForm(coeff u v), FacetBasis) on bilinear quad facets given by
    their cyclic corner nodes (mesh/task_heat.py:110-133)."""
    g, w = fem.gauss_unit(2)
    n = p.shape[1]
    x = p[:, cyc]
    rows, cols, data = [], [], []
    for r, wr in zip(g, w):
        for s, ws in zip(g, w):
            N = np.array([(1 - r) * (1 - s), r * (1 - s), r * s, (1 - r) * s])
            dNr = np.array([-(1 - s), (1 - s), s, -s])
            dNs = np.array([-(1 - r), -r, r, (1 - r)])
            tr = np.einsum("daf,a->df", x, dNr)
            ts = np.einsum("daf,a->df", x, dNs)
            jac = np.linalg.norm(np.cross(tr, ts, axis=0), axis=0) * wr * ws
            for a in range(4):
                for b in range(4):
                    rows.append(cyc[a])
                    cols.append(cyc[b])
                    data.append(coeff * N[a] * N[b] * jac)
    return sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(n, n)).tocsr()


def smoke_task_inputs(h):
    """The reference's heat smoke task (tests/test_global_flow.py:53-103) on the
    8 x 8 x 1 plate: Robin facets (h = 4e-5, T_env = 300) on the x = 0 and y = 8
    faces, one Dirichlet patch on x in [7.6, 8], y in [0, 0.4] (boundary facets
    only, as ``from_mesh_tags`` tags them).  Returns (p, t, robin matrices,
    robin loads, Dirichlet nodes)."""
    from . import mesh as omesh
    p, t = omesh.box_hex(8.0, 8.0, 1.0, h)
    srt, cyc = omesh.hex_facets(t)
    allf = np.sort(np.hstack([t[list(f)] for f in omesh._HEX_FACES]).astype(np.int64), axis=0)
    _, cnt = np.unique(allf, axis=1, return_counts=True)
    mid = p[:, srt].mean(axis=1)
    on_bnd = cnt == 1
    Bs, fs = [], []
    for sel in (omesh.in_box(mid, (0.0, 0.0), (0.0, 8.0), (0.0, 1.0)),
                omesh.in_box(mid, (0.0, 8.0), (8.0, 8.0), (0.0, 1.0))):
        ids = np.nonzero(sel & on_bnd)[0]
        Bs.append(quad_facet_mass(p, cyc[:, ids], 4.0e-5))
        f, _ = omesh.quad_facet_load(p, cyc[:, ids], 4.0e-5 * 300.0)
        fs.append(f)
    dsel = np.nonzero(omesh.in_box(mid, (7.6, 8.0), (0.0, 0.4), (0.0, 1.0)) & on_bnd)[0]
    return p, t, Bs, fs, np.unique(srt[:, dsel])
