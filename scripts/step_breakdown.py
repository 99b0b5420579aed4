"""Per-section wall-clock breakdown of optimiser iterations at C2 (device
synchronised at section exits).  Usage: python scripts/step_breakdown.py [h | c3 | c4] [iters] [logmoc | oc]"""
import json, os, sys, tempfile, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt
from sktopt._b200 import device as dev

# under torchrun: one rank per GPU, sharded run (rank 0 prints)
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
named = len(sys.argv) > 1 and sys.argv[1] in ("c3", "c4")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else ("oc" if named else "logmoc")
if named:
    sys.path.insert(0, ROOT)
    from scripts import workloads
    tsk = workloads.c3_task(sktopt) if sys.argv[1] == "c3" else workloads.c4_task(sktopt)
else:
    h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
    tsk = sktopt.mesh.toy_problem.toy_base(h)
tmp = tempfile.mkdtemp()
if named:
    cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=50, record_times=50,
                                solver_option="cg_pyamg")
    opt = sktopt.core.OC_Optimizer(cfg, tsk)
elif kind == "logmoc":
    cfg = sktopt.core.LogMOC_Config(dst_path=tmp, max_iters=200, record_times=20,
                                    vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3),
                                    solver_option="cg_pyamg")
    opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
else:
    cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=200, record_times=20, solver_option="cg_pyamg",
                                vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3))
    opt = sktopt.core.OC_Optimizer(cfg, tsk)
opt.parameterize()
opt.export_enabled = False
opt.optimize_steps(3)
opt.timer.reset()
opt.timer.cuda_sync = True
torch.cuda.synchronize()
t0 = time.perf_counter()
per_step = []
for _ in range(n):
    ts = time.perf_counter()
    opt.optimize_steps(1)
    torch.cuda.synchronize()
    per_step.append(round(1e3 * (time.perf_counter() - ts), 2))
dt = (time.perf_counter() - t0) / n
if rank != 0:
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    sys.exit(0)
print("per-step ms", per_step)
print("ms/step", dt * 1e3)
for s in opt.timer.summary():
    print(f"{s.name:60s} {s.total/n*1e3:9.2f} ms/step  n={s.count//n}")
print("precond", opt.fem.engine.precond, "pcg", opt.fem.engine.pcg_log[-2 * n:])
try:
    si = opt.filter._dev_state.solve_iters
    print("filter iters", si[-8:], "solves", len(si), "total iterations", sum(si))
except Exception as e:
    print("filter iters: n/a", type(e).__name__)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
