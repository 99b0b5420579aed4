#!/usr/bin/env python
"""Generates tests/golden/reference_*.npz from the REAL reference (scikit-topt
0.3.9 + scikit-fem + scipy), for the day an environment with those packages is
reachable.  It cannot run in the build container or on the GPU box (scikit-fem,
pyamg, meshio are not installed and there is no network), so parity is pinned to
the reference only once somebody runs:

    pip install scikit-fem pyamg meshio matplotlib imageio numba
    PYTHONPATH=/root/reference/scikit-topt python tests/golden/make_reference_fixtures.py

and commits the three .npz files; tests/test_reference_fixtures.py then compares
the oracle (CPU) and the CUDA path (GPU) with them and skips LOUDLY until then.

Per task (toy_test: 192-hex cantilever; toy2: two load cases; heat: the
reference's own heat smoke task of tests/test_global_flow.py:53-103) it stores:
mesh p / t, Dirichlet dofs, loads, design / fixed elements, element volumes; the
assembled K(rho) of a seeded density (CSR arrays); the displacement / temperature
field, compliance and element energies of that density; Helmholtz and spatial
filter forward / gradient of seeded vectors; and the per-iteration compliance and
final density of 5 OC and 5 LogMOC iterations."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _csr(K):
    K = K.tocsr()
    K.sort_indices()
    return dict(indptr=K.indptr, indices=K.indices, data=K.data)


def _run(sktopt, kind, tsk, iters, **cfg_kw):
    Cfg = sktopt.core.OC_Config if kind == "oc" else sktopt.core.LogMOC_Config
    Opt = sktopt.core.OC_Optimizer if kind == "oc" else sktopt.core.LogMOC_Optimizer
    with tempfile.TemporaryDirectory() as tmp:
        cfg = Cfg(dst_path=tmp, max_iters=iters, record_times=iters, **cfg_kw)
        opt = Opt(cfg, tsk)
        opt.parameterize()
        opt.optimize()
        rec = opt.recorder.as_object()
        name = "compliance" if hasattr(rec, "compliance") else str(getattr(tsk, "objective", ""))
        hist = np.asarray(getattr(rec, name))
        st = getattr(opt, "_state", None)
        rho = np.asarray(st.rho if st is not None else opt.rho)
    return hist, rho


def elasticity_fixture(sktopt, make_task, name):
    from sktopt.fea import composer, solver
    from sktopt import filters
    tsk = make_task()
    tsk.exlude_dirichlet_from_design()
    mesh, basis = tsk.mesh, tsk.basis
    rng = np.random.default_rng(0)
    rho = rng.uniform(0.05, 1.0, mesh.nelements)
    out = dict(p=mesh.p, t=mesh.t, dirichlet_dofs=np.asarray(tsk.dirichlet_dofs),
               design=np.asarray(tsk.design_elements), fixed=np.asarray(tsk.fixed_elements),
               volumes=np.asarray(tsk.elements_volume), rho=rho, E=tsk.E, nu=tsk.nu)
    forces = tsk.neumann_linear if isinstance(tsk.neumann_linear, list) else [tsk.neumann_linear]
    out["forces"] = np.stack(forces)
    K = composer.assemble_stiffness_matrix(basis, rho, tsk.E, tsk.E * 1e-3, 3.0, tsk.nu)
    for k, v in _csr(K).items():
        out["K_" + k] = v
    fem = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="spsolve")
    u = np.zeros((basis.N, len(forces)))
    out["compliance"] = np.asarray(fem.objectives_multi_load(rho, 3.0, u))
    out["u"] = u
    out["energy"] = np.asarray(fem.energy_multi_load(rho, 3.0, u))
    v = rng.uniform(-1.0, 0.0, mesh.nelements)
    for fname, F in (("helmholtz", filters.HelmholtzFilterNodal), ("spatial", filters.SpacialFilter)):
        f = F.from_defaults(mesh, tsk.elements_volume, 0.6 if fname == "helmholtz" else 1.5,
                            design_mask=tsk.design_mask)
        out[fname + "_forward"] = np.asarray(f.forward(rho))
        out[fname + "_gradient"] = np.asarray(f.gradient(v))
    out["filter_input_v"] = v
    for kind in ("oc", "logmoc"):
        hist, rho_fin = _run(sktopt, kind, make_task(), 5)
        out[kind + "_history"] = hist
        out[kind + "_rho_final"] = rho_fin
    np.savez_compressed(os.path.join(HERE, f"reference_{name}.npz"), **out)
    print("wrote reference_%s.npz" % name)


def heat_task(sktopt, mesh_size=0.5, intorder=2):
    import skfem
    x_len, y_len, z_len = 8.0, 8.0, 1.0
    mesh = sktopt.mesh.toy_problem.create_box_hex(x_len, y_len, z_len, mesh_size)
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({
        "robin_0": rng((0.0, 0.0), (0.0, y_len), (0.0, z_len)),
        "robin_1": rng((0.0, x_len), (y_len, y_len), (0.0, z_len)),
        "dirichlet_0": rng((x_len - 1.0 * x_len / 20, x_len), (0.0, 1.0 * y_len / 20), (0.0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.array(range(mesh.nelements))})
    basis = skfem.Basis(mesh, skfem.ElementHex1(), intorder=intorder)
    return sktopt.mesh.LinearHeatConduction.from_mesh_tags(
        basis, 600.0, 4.0e-5, 300.0, True, 10.0, "compliance")


def heat_fixture(sktopt):
    out = {}
    for intorder in (1, 2):
        tsk = heat_task(sktopt, 0.5, intorder)
        rho = np.random.default_rng(0).uniform(0.1, 0.95, tsk.mesh.nelements)
        fem = sktopt.fea.FEM_SimpLinearHeatConduction(tsk, 1e-3)
        T = np.zeros((tsk.basis.N, 1))
        J = fem.objectives_multi_load(rho, 3.0, T)
        g = fem.compliance_sensitivity_multi_load(rho, 3.0, T)
        tag = "io%d_" % intorder
        out.update({tag + "rho": rho, tag + "J": np.asarray(J), tag + "T": T,
                    tag + "sensitivity": np.asarray(g),
                    tag + "energy": np.asarray(fem.energy_multi_load(rho, 3.0, T))})
        if intorder == 2:
            out["p"], out["t"] = tsk.mesh.p, tsk.mesh.t
            hist, rho_fin = _run(sktopt, "oc", heat_task(sktopt, 0.5, 2), 5)
            out["oc_history"], out["oc_rho_final"] = hist, rho_fin
    np.savez_compressed(os.path.join(HERE, "reference_heat.npz"), **out)
    print("wrote reference_heat.npz")


def main():
    try:
        import skfem  # noqa: F401
        import sktopt
    except ImportError as e:
        print("the reference cannot be imported here (%s): nothing written" % e)
        return 1
    if "_b200" in getattr(sktopt, "__file__", "") or hasattr(sktopt, "_b200"):
        print("this is the B200 build of `sktopt`, not the reference: put the reference's "
              "scikit-topt/ directory first on PYTHONPATH")
        return 1
    elasticity_fixture(sktopt, sktopt.mesh.toy_problem.toy_test, "toy_test")
    elasticity_fixture(sktopt, sktopt.mesh.toy_problem.toy2, "toy2")
    heat_fixture(sktopt)
    return 0


if __name__ == "__main__":
    sys.exit(main())
