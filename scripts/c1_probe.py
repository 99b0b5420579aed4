"""BASELINE config 1 (52,728 hex, 50 OC iterations) against the oracle fixture for several
filter / state solve tolerances: how the optimiser amplifies the solve errors."""
import os, sys, tempfile, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200")); sys.path.insert(0, ROOT)
import sktopt
from sktopt.fea.solver_elastic import LinearSolverConfig
from sktopt.filters import helmholtz_filter_nodal as hf
ref = np.load(os.path.join(ROOT, "tests", "golden", "c1_oc50_oracle.npz"))
for frtol, srtol in ((1e-12, 1e-8), (3e-13, 1e-8), (1e-13, 1e-8)):
    hf._HelmholtzDevice.RTOL = frtol
    tsk = sktopt.mesh.toy_problem.toy_base(float(ref["mesh_size"]))
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=50, record_times=50, solver_option="cg_pyamg")
        cfg.solver_config = LinearSolverConfig(solver="cg_pyamg", rtol=srtol)
        opt = sktopt.core.OC_Optimizer(cfg, tsk); opt.parameterize(); opt.export_enabled = False
        opt.optimize()
        comp = np.asarray(opt.recorder.as_object().compliance); verr = np.asarray(opt.recorder.as_object().vol_error)
        rho = opt._state.rho.cpu().numpy()
    rel = np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])
    print("filter rtol %g state rtol %g: compliance %.2e (it1 %.1e) drho %.2e verr %.2e steps_equal %s nonconv %d" % (
        frtol, srtol, rel.max(), rel[0], np.abs(rho - ref["rho_final"]).max(), np.abs(verr - ref["vol_error"]).max(),
        list(opt.bisection_steps) == [int(v) for v in ref["bisection_steps"]], sum(1 for l in opt.fem.engine.pcg_log if not l[1])), flush=True)
