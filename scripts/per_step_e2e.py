"""Per-step wall time of the e2e loop of bench.py (H2D of rho, step, D2H, sync)."""
import os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 45
tsk = sktopt.mesh.toy_problem.toy_base(0.0577)
cfg = sktopt.core.LogMOC_Config(dst_path=tempfile.mkdtemp(), max_iters=200, record_times=20,
                                vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3),
                                solver_option="cg_pyamg")
opt = sktopt.core.LogMOC_Optimizer(cfg, tsk); opt.parameterize(); opt.export_enabled = False
opt.optimize_steps(25)
st = opt._state
ne = st.rho.numel()
rho_h = torch.empty(ne, dtype=torch.float64).pin_memory(); out_h = torch.empty(ne, dtype=torch.float64).pin_memory()
rho_h.copy_(st.rho)
torch.cuda.synchronize()
parts = []
for i in range(n):
    t0 = time.perf_counter()
    if os.environ.get('NO_H2D') != '1': st.rho.copy_(rho_h, non_blocking=True)
    t1 = time.perf_counter()
    opt.optimize_steps(1)
    t2 = time.perf_counter()
    if os.environ.get('NO_D2H') != '1': out_h.copy_(st.rho, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if os.environ.get('NO_HOSTCOPY') != '1': rho_h.copy_(out_h)
    t4 = time.perf_counter()
    parts.append((round(1e3*(t1-t0),2), round(1e3*(t2-t1),2), round(1e3*(t3-t2),2), round(1e3*(t4-t3),2)))
print("h2d / step / d2h+sync / host copy (ms):", parts[5:15])
import numpy as np
print("mean total ms", np.mean([sum(p) for p in parts[3:]]))
