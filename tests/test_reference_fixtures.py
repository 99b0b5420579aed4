"""Consumes tests/golden/reference_*.npz -- outputs of the REAL reference
(scikit-topt 0.3.9 + scikit-fem) written by tests/golden/make_reference_fixtures.py
-- when they exist: the oracle (CPU) and the CUDA path (GPU) are compared with
them at the north-star tolerances (pattern bit-exact, K and sensitivities 1e-10,
compliance 1e-6, densities 1e-4).  The reference cannot be installed in the build
container (no scikit-fem, no network), so until somebody generates and commits the
files these tests SKIP with that message: parity stays pinned to the oracle only
(DESIGN.md section 5)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
WHY = ("tests/golden/reference_{}.npz not present: generate it with the real reference "
       "(tests/golden/make_reference_fixtures.py); PARITY TO THE REFERENCE IS UNPINNED until then")


def _load(name):
    path = os.path.join(HERE, "golden", f"reference_{name}.npz")
    if not os.path.exists(path):
        pytest.skip(WHY.format(name))
    return np.load(path)


@pytest.mark.parametrize("name", ["toy_test", "toy2"])
def test_oracle_against_reference_elasticity(name):
    import scipy.sparse as sp
    from oracle import fem, filters as ofilters, optim
    ref = _load(name)
    p, t, rho = ref["p"], ref["t"], ref["rho"]
    E, nu = float(ref["E"]), float(ref["nu"])
    K = fem.assemble_stiffness(p, t, rho, E, E * 1e-3, 3.0, nu)
    assert np.array_equal(K.indptr, ref["K_indptr"]) and np.array_equal(K.indices, ref["K_indices"])
    assert np.abs(K.data - ref["K_data"]).max() <= 1e-10 * np.abs(ref["K_data"]).max()
    forces = [f for f in ref["forces"]]
    comp, U = fem.compliance_multi(p, t, rho, E, E * 1e-3, 3.0, nu, forces, ref["dirichlet_dofs"])
    assert np.abs(comp - ref["compliance"]).max() <= 1e-6 * np.abs(ref["compliance"]).max()
    en = fem.strain_energy(p, t, rho, U, E, E * 1e-3, 3.0, nu)
    assert np.abs(en - ref["energy"]).max() <= 1e-10 * np.abs(ref["energy"]).max()
    mask = np.isin(np.arange(t.shape[1]), ref["design"])
    h = ofilters.HelmholtzOracle(p, t, ref["volumes"], mask)
    h.set_radius(0.6)
    assert np.abs(h.forward(rho) - ref["helmholtz_forward"]).max() <= 1e-9
    assert np.abs(h.gradient(ref["filter_input_v"]) - ref["helmholtz_gradient"]).max() <= 1e-9
    s = ofilters.SpatialOracle(p, t, mask)
    s.set_radius(1.5)
    assert np.abs(s.forward(rho) - ref["spatial_forward"]).max() <= 1e-12
    pinned = np.setdiff1d(np.arange(t.shape[1]), np.union1d(ref["design"], ref["fixed"]))
    pr = optim.Problem(p, t, ref["dirichlet_dofs"], forces, ref["design"], pinned, ref["volumes"],
                       E, nu, fixed=ref["fixed"])
    for kind in ("oc", "logmoc"):
        out = optim.run(pr, kind, max_iters=5)
        h_ref = ref[kind + "_history"]
        assert np.abs(np.asarray(out["compliance"]) - h_ref).max() <= 1e-6 * np.abs(h_ref).max()
        assert np.abs(out["rho_final"] - ref[kind + "_rho_final"]).max() <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["toy_test", "toy2"])
def test_cuda_path_against_reference_elasticity(name):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ref = _load(name)
    import tempfile
    import sktopt
    make = getattr(sktopt.mesh.toy_problem, name)
    tsk = make()
    tsk.exlude_dirichlet_from_design()
    assert np.array_equal(tsk.mesh.t, ref["t"]) and np.array_equal(tsk.mesh.p, ref["p"])
    assert np.array_equal(np.asarray(tsk.dirichlet_dofs), ref["dirichlet_dofs"])
    rho = ref["rho"]
    K = sktopt.fea.composer.assemble_stiffness_matrix(tsk.basis, rho, tsk.E, tsk.E * 1e-3, 3.0,
                                                     tsk.nu)
    assert np.array_equal(K.indptr, ref["K_indptr"]) and np.array_equal(K.indices, ref["K_indices"])
    assert np.abs(K.data - ref["K_data"]).max() <= 1e-10 * np.abs(ref["K_data"]).max()
    fem = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="spsolve")
    u = np.zeros_like(ref["u"])
    comp = fem.objectives_multi_load(rho, 3.0, u)
    assert np.abs(comp - ref["compliance"]).max() <= 1e-6 * np.abs(ref["compliance"]).max()
    en = fem.energy_multi_load(rho, 3.0, u)
    assert np.abs(en - ref["energy"]).max() <= 1e-6 * np.abs(ref["energy"]).max()
    for kind, Cfg, Opt in (("oc", sktopt.core.OC_Config, sktopt.core.OC_Optimizer),
                           ("logmoc", sktopt.core.LogMOC_Config, sktopt.core.LogMOC_Optimizer)):
        with tempfile.TemporaryDirectory() as tmp:
            opt = Opt(Cfg(dst_path=tmp, max_iters=5, record_times=5), make())
            opt.parameterize()
            opt.optimize()
            hist = np.asarray(opt.recorder.as_object().compliance)
            rho_fin = opt._state.rho.cpu().numpy()
        h_ref = ref[kind + "_history"]
        assert np.abs(hist - h_ref).max() <= 1e-6 * np.abs(h_ref).max()
        assert np.abs(rho_fin - ref[kind + "_rho_final"]).max() <= 1e-4


def test_oracle_against_reference_heat():
    ref = _load("heat")
    assert "io2_J" in ref.files      # consumed by tests/test_gpu_heat.py once present
