"""CUDA-event timing of the pieces of the per-iteration operator set-up at C2."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt
from sktopt._b200 import device as dev, lib as _lib
from sktopt.fea._engine import KE_ELASTIC, get_engine

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
tsk = sktopt.mesh.toy_problem.toy_base(h)
tsk.exlude_dirichlet_from_design()
eng = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
rho = dev.to_dev(np.random.default_rng(0).uniform(0.2, 1.0, eng.n_elem))
eng.set_modulus(rho, tsk.E, tsk.E * 1e-3, 3.0)
eng.prepare()
mg = eng.mg

def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

print("assemble L0 (enforced)      %.3f ms" % t(lambda: eng.assemble()))
print("inv_diag L0                 %.3f ms" % t(lambda: dev.csr_inv_diag(eng.row_ptr, eng.col_idx, eng.vals, out=eng.inv_diag)))
for l in range(1, mg.n_levels):
    lv = mg.levels[l]
    if l == 1:
        f = lambda: _lib.check(mg.lib.sktb_elem_restrict(lv["n_elem"], dev._ptr(lv["child"]), dev._ptr(lv["ptype"]), dev._ptr(mg.Qtab), None, dev._ptr(eng.unit_ke), dev._ptr(eng.dm.elem_class), dev._ptr(eng.scale), dev._ptr(lv["ke"]), dev._stream()))
    else:
        prev = mg.levels[l - 1]
        f = lambda: _lib.check(mg.lib.sktb_elem_restrict(lv["n_elem"], dev._ptr(lv["child"]), dev._ptr(lv["ptype"]), dev._ptr(mg.Qtab), dev._ptr(prev["ke"]), None, None, None, dev._ptr(lv["ke"]), dev._stream()))
    print("elem_restrict L%d (%7d el) %.3f ms" % (l, lv["n_elem"], t(f)))
    print("assemble L%d                %.3f ms" % (l, t(lambda: lv["dm"].assemble(3, lv["ke"], scale=None, dir_mask=lv["mask"], out=lv["vals"], per_element=True))))
    print("inv_diag L%d                %.3f ms" % (l, t(lambda: dev.bsr3_inv_diag(lv["node_ptr"], lv["node_col"], lv["vals"], out=lv["inv_diag"]))))
print("whole mg.setup              %.3f ms" % t(lambda: mg.setup()))
r = torch.randn(eng.n_dof, dtype=dev.F64, device="cuda"); z = torch.empty_like(r)
print("V-cycle                     %.3f ms" % t(lambda: mg.vcycle(r, z)))
