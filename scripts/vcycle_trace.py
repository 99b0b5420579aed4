"""One steady-state solve at C2 inside a cudaProfilerStart/Stop window (for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt
from sktopt._b200 import device as dev
from sktopt.fea._engine import KE_ELASTIC, get_engine

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
what = sys.argv[2] if len(sys.argv) > 2 else "vcycle"
tsk = sktopt.mesh.toy_problem.toy_base(h)
tsk.exlude_dirichlet_from_design()
eng = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
rho = dev.to_dev(np.random.default_rng(0).uniform(0.2, 1.0, eng.n_elem))
eng.set_modulus(rho, tsk.E, tsk.E * 1e-3, 3.0)
eng.prepare()
r = torch.randn(eng.n_dof, dtype=dev.F64, device="cuda"); z = torch.empty_like(r)
eng.mg.vcycle(r, z)
torch.cuda.synchronize()
torch.cuda.profiler.start()
if what == "vcycle":
    eng.mg.vcycle(r, z)
elif what == "setup":
    eng.prepare()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
