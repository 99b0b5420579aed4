"""C2 assembled operator + a few launches of the node-block TMA SpMV, nothing else
(target of `ncu --set full -k regex:spmv_bsr3_tma -s 1 -c 1`: DRAM traffic per launch)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
os.environ["SKTOPT_B200_MATFREE"] = "0"
os.environ["SKTOPT_B200_PRECOND"] = "jacobi"
import sktopt
from sktopt._b200 import device as dev
from sktopt.fea._engine import KE_ELASTIC, get_engine
tsk = sktopt.mesh.toy_problem.toy_base(float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577)
tsk.exlude_dirichlet_from_design()
eng = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
eng.set_modulus(dev.to_dev(np.random.default_rng(0).uniform(0.2, 1.0, eng.n_elem)), tsk.E, tsk.E * 1e-3, 3.0)
eng.assemble(enforce=True)
x = torch.randn(eng.n_dof, dtype=dev.F64, device="cuda"); y = torch.empty_like(x)
for _ in range(4):
    dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, x, eng.max_deg, out=y)
torch.cuda.synchronize()
print("nnz", eng.nnz)
