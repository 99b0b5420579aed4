"""Matrix-free grid operator vs assembled TMA SpMV at C2 (CUDA events, L2 flushed by size)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt
from sktopt._b200 import device as dev
from sktopt.fea._engine import KE_ELASTIC, get_engine

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
tsk = sktopt.mesh.toy_problem.toy_base(h)
tsk.exlude_dirichlet_from_design()
eng = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
print("matrix_free", eng.matrix_free, "n_nodes", eng.dm.n_nodes, "precond", eng.precond)
rho = dev.to_dev(np.random.default_rng(0).uniform(0.2, 1.0, eng.n_elem))
eng.set_modulus(rho, tsk.E, tsk.E * 1e-3, 3.0)
eng.prepare()

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

x = torch.randn(eng.n_dof, dtype=dev.F64, device="cuda"); y = torch.empty_like(x)
ms = t(lambda: eng.gridop.apply(x, out=y))
nn = eng.dm.n_nodes
print("gridop apply   %.4f ms  (%.1f GFLOP/s fp64, %.1f GB/s x+y+E)" % (ms, nn * 1200 / ms / 1e6, (nn * 48 + eng.n_elem * 8) / ms / 1e6))
print("gridop invdiag %.4f ms" % t(lambda: eng.gridop.inv_diag(out=eng.inv_diag)))
print("mg.setup       %.4f ms" % t(lambda: eng.mg.setup(), 5))
r = torch.randn(eng.n_dof, dtype=dev.F64, device="cuda"); z = torch.empty_like(r)
print("V-cycle        %.4f ms" % t(lambda: eng.mg.vcycle(r, z)))
if len(sys.argv) > 2:
    eng.assemble(enforce=True)
    ms2 = t(lambda: dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, x, eng.max_deg, out=y))
    y2 = eng.gridop.apply(x)
    print("assembled TMA  %.4f ms   max|diff| %.3e" % (ms2, float((y - y2).abs().max() / y.abs().max())))
