import sys, numpy as np, torch
sys.path.insert(0, "scikit-topt_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import sktopt
from sktopt._b200 import device as dev
import test_gpu_long as tl
tsk = tl._tet_task(sktopt)
tsk.exlude_dirichlet_from_design()
print("nelem", tsk.mesh.nelements, "nodes", tsk.mesh.nvertices, "vol min", tsk.elements_volume.min())
fem = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3)
eng = fem.engine
print("max_deg", eng.max_deg, "n_class", eng.dm.n_class, "format", eng.spmv_format)
rho = np.full(tsk.mesh.nelements, 0.5)
u = np.zeros((tsk.basis.N, 2))
c = fem.objectives_multi_load(rho, 1.0, u)
print("compliance", c, "pcg", eng.pcg_log, "u nan", np.isnan(u).any())
print("vals nan", torch.isnan(eng.vals).any().item(), "minv nan", torch.isnan(eng.inv_diag).any().item(), "rhs nan", torch.isnan(eng.rhs).any().item())
E = fem.energy_multi_load(rho, 1.0, u)
print("energy nan", np.isnan(E).any(), E.sum(axis=0))
f = sktopt.filters.HelmholtzFilterNodal.from_defaults(tsk.mesh, tsk.elements_volume, 0.5, design_mask=tsk.design_mask)
r = f.forward(rho); print("filter fwd nan", np.isnan(r).any(), r.min(), r.max())
g = f.gradient(-np.abs(E[:,0])); print("filter grad nan", np.isnan(g).any())
