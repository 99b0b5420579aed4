"""Per-step device time of a C2 LogMOC run (CUDA events per step)."""
import os, sys, tempfile
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
ts = []
for i in range(n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); opt.optimize_steps(1); b.record(); torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b), 2))
print("ms:", ts)
print("pcg:", [l[0] for l in opt.fem.engine.pcg_log])
