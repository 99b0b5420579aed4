"""One steady-state optimiser iteration at C2 inside a cudaProfilerStart/Stop window (for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`)."""
import os, sys, tempfile
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt

h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
tsk = sktopt.mesh.toy_problem.toy_base(h)
cfg = sktopt.core.LogMOC_Config(dst_path=tempfile.mkdtemp(), max_iters=200, record_times=20,
                                vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3),
                                solver_option="cg_pyamg")
opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
opt.parameterize()
opt.export_enabled = False
opt.optimize_steps(3)
torch.cuda.synchronize()
torch.cuda.profiler.start()
opt.optimize_steps(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
