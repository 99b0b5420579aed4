#!/usr/bin/env python
"""Generates the committed fixtures in tests/golden/ (run from the repo root:
``python tests/golden/make_golden.py``).

The reference (scikit-topt 0.3.9) cannot be imported in this image -- scikit-fem
and pyamg are not installed and there is no network -- so there
are NO reference-generated vectors here.  Two kinds of fixtures instead:

``exact_element_matrices.npz``
    Element matrices integrated EXACTLY with sympy rationals (no quadrature, no
    NumPy): the elasticity, Laplace and mass matrices of a trilinear brick
    (1/2 x 1/4 x 1/5, nu = 3/10, E = 1, local vertex order = skfem ElementHex1,
    SURVEY.md App. A.1) and of an affine P1 tetrahedron.  They pin the oracle's
    and the CUDA path's quadrature / Jacobian / Voigt-free bilinear form
    (fea/composer.py:82-98,139-141; filters/helmholtz_filter_nodal.py:136-141).

``oracle_regression.npz``
    Outputs of the oracle (oracle/*.py) on small seeded cases of every stage of
    the path.  They are REGRESSION vectors: they freeze the oracle the CUDA path
    was validated against, so that later edits to either side show up; they do
    not pin the oracle to the reference ("parity unpinned").
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# skfem ElementHex1 local vertex order on the reference cube
HEX_VERTS = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0),
             (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)]
BRICK = ("1/2", "1/4", "1/5")
TET = [(0, 0, 0), ("1/2", 0, 0), ("1/10", "1/3", 0), ("1/7", "1/9", "1/4")]
NU = "3/10"


def exact_matrices():
    import sympy as sy
    x, y, z = sy.symbols("x y z")
    nu = sy.Rational(NU)
    lam = nu / ((1 + nu) * (1 - 2 * nu))
    mu = sy.Rational(1, 2) / (1 + nu)

    def forms(N, integrate):
        n = len(N)
        g = [[sy.diff(Na, v) for v in (x, y, z)] for Na in N]
        Ke = sy.zeros(3 * n, 3 * n)
        Le = sy.zeros(n, n)
        Me = sy.zeros(n, n)
        for a in range(n):
            for b in range(a, n):
                gg = [[integrate(g[a][i] * g[b][j]) for j in range(3)] for i in range(3)]
                tr = gg[0][0] + gg[1][1] + gg[2][2]
                Le[a, b] = Le[b, a] = tr
                Me[a, b] = Me[b, a] = integrate(N[a] * N[b])
                for i in range(3):
                    for j in range(3):
                        # lam div(u) div(v) + 2 mu e(u):e(v); row = test dof 3a+i,
                        # col = trial dof 3b+j (u = N_b e_j, v = N_a e_i)
                        val = lam * gg[i][j] + mu * gg[j][i] + (mu * tr if i == j else 0)
                        Ke[3 * a + i, 3 * b + j] = val
                        Ke[3 * b + j, 3 * a + i] = val
        f = lambda M: np.array(M.evalf(40).tolist(), dtype=np.float64)
        return f(Ke), f(Le), f(Me)

    a, b, c = (sy.Rational(s) for s in BRICK)
    Nh = []
    for vx, vy, vz in HEX_VERTS:
        fx = x / a if vx else 1 - x / a
        fy = y / b if vy else 1 - y / b
        fz = z / c if vz else 1 - z / c
        Nh.append(fx * fy * fz)
    hex_int = lambda e: sy.integrate(sy.expand(e), (x, 0, a), (y, 0, b), (z, 0, c))
    hexK, hexL, hexM = forms(Nh, hex_int)

    P = sy.Matrix([[sy.Rational(v) for v in row] for row in TET])
    A = sy.Matrix([[1, *P.row(k)] for k in range(4)])
    coef = A.inv()                      # N_k = coef[0,k] + coef[1,k] x + ...
    Nt = [coef[0, k] + coef[1, k] * x + coef[2, k] * y + coef[3, k] * z for k in range(4)]
    vol = abs(A.det()) / 6
    # integrate polynomials (degree <= 2) over the tetrahedron exactly through
    # the affine map from the unit tetrahedron
    r, s, t = sy.symbols("r s t")
    J = sy.Matrix([[P[k, d] - P[0, d] for k in (1, 2, 3)] for d in range(3)])
    xm = sy.Matrix([P[0, d] for d in range(3)]) + J * sy.Matrix([r, s, t])
    detJ = abs(J.det())

    def tet_int(e):
        e = sy.expand(e.subs({x: xm[0], y: xm[1], z: xm[2]}, simultaneous=True))
        return detJ * sy.integrate(e, (t, 0, 1 - r - s), (s, 0, 1 - r), (r, 0, 1))
    tetK, tetL, tetM = forms(Nt, tet_int)
    # unit cube (the toy meshes are grids of cubes: Ke(h) = h Ke(1), Le(h) = h Le(1),
    # Me(h) = h^3 Me(1))
    Nc = []
    for vx, vy, vz in HEX_VERTS:
        Nc.append((x if vx else 1 - x) * (y if vy else 1 - y) * (z if vz else 1 - z))
    cube_int = lambda e: sy.integrate(sy.expand(e), (x, 0, 1), (y, 0, 1), (z, 0, 1))
    cubeK, cubeL, cubeM = forms(Nc, cube_int)
    hex_p = np.array([[float(sy.Rational(BRICK[d])) * v[d] for v in HEX_VERTS]
                      for d in range(3)])
    tet_p = np.array([[float(sy.Rational(v)) for v in row] for row in TET]).T
    return dict(nu=float(nu), cube_K=cubeK, cube_L=cubeL, cube_M=cubeM,
                hex_p=hex_p, hex_K=hexK, hex_L=hexL, hex_M=hexM,
                tet_p=tet_p, tet_K=tetK, tet_L=tetL, tet_M=tetM, tet_vol=float(vol))


def oracle_regression():
    from oracle import fem, filters as ofilters, heat as oheat, mesh as omesh, optim
    out = {}
    o = omesh.toy_base(1.0)                       # the reference's toy_test(): 192 hex
    p, t = o["p"], o["t"]
    ne = t.shape[1]
    rng = np.random.default_rng(20261017)
    rho = rng.uniform(0.2, 1.0, ne)
    E0, Emin, pw, nu = o["E"], o["E"] * 1e-3, 3.0, o["nu"]
    K = fem.assemble_stiffness(p, t, rho, E0, Emin, pw, nu, 2)
    K.sort_indices()
    out["toy_rho"] = rho
    out["toy_K_indptr"], out["toy_K_indices"], out["toy_K_data"] = K.indptr, K.indices, K.data
    C, u = fem.compliance_single(p, t, rho, E0, Emin, pw, nu, o["force"], o["dirichlet_dofs"],
                                 solver="spsolve")[:2]
    out["toy_compliance"], out["toy_u"] = C, u
    U = fem.strain_energy(p, t, rho, u, E0, Emin, pw, nu, 2)[:, 0]
    out["toy_energy"] = U
    out["toy_dC"] = optim.dC_drho_simp(rho, U, E0, Emin, pw)
    dmask = np.isin(np.arange(ne), o["design"])
    v = rng.standard_normal(ne)
    out["toy_v"] = v
    hf = ofilters.HelmholtzOracle(p, t, o["volumes"], dmask)
    hf.set_radius(0.6)
    out["toy_helmholtz_forward"], out["toy_helmholtz_gradient"] = hf.forward(rho), hf.gradient(v)
    sf = ofilters.SpatialOracle(p, t, dmask)
    sf.set_radius(1.5)
    out["toy_spatial_forward"], out["toy_spatial_gradient"] = sf.forward(rho), sf.gradient(v)
    out["toy_heaviside"] = optim.heaviside(rho, 4.0, 0.5)
    out["toy_heaviside_derivative"] = optim.heaviside_derivative(rho, 4.0, 0.5)
    pr = optim.Problem(p, t, o["dirichlet_dofs"], o["force"], o["design"], o["pinned"],
                       o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    for method, kw in (("oc", dict(vol_frac=0.6)), ("logmoc", dict(vol_frac=0.6))):
        h = optim.run(pr, method, max_iters=8, filter_radius=0.6, solver="spsolve", **kw)
        out[f"toy_{method}_compliance"] = np.array(h["compliance"])
        out[f"toy_{method}_vol_error"] = np.array(h["vol_error"])
        out[f"toy_{method}_rho_final"] = h["rho_final"]
    # heat smoke task of the reference (tests/test_global_flow.py:53-103), h = 0.5
    hp, ht, Bs, fs, D = oheat.smoke_task_inputs(0.5)
    hrho = rng.uniform(0.1, 0.95, ht.shape[1])
    out["heat_rho"] = hrho
    for obj, w in (("compliance", 0.0), ("heat_exchange", 0.0), ("heat_exchange", 0.25),
                   ("averaged_temp", 0.0)):
        J, T, lam, _ = oheat.objectives(hp, ht, hrho, 10.0, 1e-2, 3.0, 4, 4.0e-5, 300.0,
                                        Bs, fs, D, 600.0, obj, 2, w)
        tag = f"heat_{obj}_{w:g}"
        out[tag + "_J"], out[tag + "_T"], out[tag + "_lambda"] = J, T, lam
        if obj != "compliance":
            out[tag + "_energy"] = oheat.grad_dot_energy(hp, ht, T, lam, 2)
    g, Uh = oheat.sensitivity(hp, ht, hrho, out["heat_compliance_0_T"], 10.0, 1e-2, 3.0, 4,
                              4.0e-5, 300.0, 2)
    out["heat_compliance_sensitivity"], out["heat_compliance_energy"] = g, Uh
    return out


def main():
    ex = exact_matrices()
    np.savez_compressed(os.path.join(HERE, "exact_element_matrices.npz"), **ex)
    if "--exact-only" in sys.argv:        # keep the frozen regression vectors
        print("exact_element_matrices.npz",
              os.path.getsize(os.path.join(HERE, "exact_element_matrices.npz")), "bytes")
        return
    reg = oracle_regression()
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **reg)
    for f in ("exact_element_matrices.npz", "oracle_regression.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
