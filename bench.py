#!/usr/bin/env python
"""Headline benchmark: optimiser iterations / second on the 1M-element 3D
cantilever (BASELINE.json configs[1]: log-space MOC, vol frac 0.3, one B200),
plus the achieved HBM bandwidth of the PCG SpMV against the measured roofline.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one full optimiser iteration (filter -> assemble -> PCG solve ->
element energy -> sensitivity -> filter adjoint -> LogMOC update) with export
ticks switched off.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "optimizer iters/sec on 1M-elem 3D cantilever; PCG SpMV HBM GB/s vs peak"
UNIT = "iters/s"
C2_MESH_SIZE = 0.0577          # toy_base(0.0577): 139x104x70 = 1,011,920 hex
C5_MESH_SIZE = 0.0288          # toy_base(0.0288): 278x209x139 = 8,076,178 hex (BASELINE configs[4])



_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries below us write there too
    (NCCL prints its version banner on fd 1 when the box sets NCCL_DEBUG), so fd 1
    is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

def workload_name(mesh_size: float) -> str:
    tag = ("C2" if abs(mesh_size - C2_MESH_SIZE) < 1e-12 else
           "C5" if abs(mesh_size - C5_MESH_SIZE) < 1e-12 else
           "C1" if abs(mesh_size - C1_MESH_SIZE) < 1e-12 else "custom")
    return "%s: 3D cantilever toy_base(%g), LogMOC, vol_frac 0.3" % (tag, mesh_size)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------- CPU legs --
# The CPU side is the oracle's restatement of the reference path with its heavy
# stages in C / OpenMP (oracle/cport: assembly, scipy-cg-semantics Jacobi PCG,
# element energies) so that it runs the SAME mesh as the GPU arm on all host
# cores -- nothing is extrapolated from a smaller mesh.  The Helmholtz systems
# are solved by scipy cg (a sparse LU of a 1M-node 3-D system does not fit).
C1_MESH_SIZE = 0.155           # toy_base(0.155): 52 x 39 x 26 = 52,728 hex (BASELINE configs[0])
CPU_C2_BUDGET_S = 150.0        # stop starting new C2 steps after this long (>= 1 step always)
CPU_C2_MAX_STEPS = 2
C1_STEPS, C1_WARMUP = 3, 1


def cpu_port_run(mesh_size, method, max_iters, n_warm, n_timed, vol_frac, budget_s=None):
    """Full optimiser iterations of the CPU port on toy_base(mesh_size).
    Returns a dict with the measured seconds per step (mean over the timed
    steps actually run), the step count, CG iterations and the thread count."""
    from oracle import cport, mesh as omesh, optim
    # all the host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to
    # its workers, but in the reference arm rank 0 is the only rank that works
    n_thr = int(os.environ.get("SKTOPT_BENCH_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    cport.lib().cport_set_threads(n_thr)
    t0 = time.perf_counter()
    o = omesh.toy_base(mesh_size)
    pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                       o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    be = cport.CBackend(o["p"], o["t"], 3, cport.unit_elasticity_ke(o["p"], o["t"], o["nu"]),
                        o["dirichlet_dofs"])
    mats = cport.scalar_matrices(o["p"], o["t"])
    setup_s = time.perf_counter() - t0
    marks, tm = [], {}
    res = optim.run(pr, method, max_iters=max_iters, iters=n_warm + n_timed, vol_frac=vol_frac,
                    solver="cg_jacobi", rtol=1e-8, cg_maxiter=200000, backend=be,
                    filter_solver="cg", filter_matrices=mats, step_times=marks, timings=tm,
                    time_budget=budget_s)
    done = len(marks) - 1
    warm = min(n_warm, max(done - 1, 0))
    timed = done - warm
    dt = (marks[-1] - marks[warm]) / timed
    return dict(s_per_step=dt, steps_timed=timed, warmup=warm, n_elem=int(o["t"].shape[1]),
                cg_iters=[int(v) for v in res["cg_iters"]], threads=cport.num_threads(),
                setup_s=setup_s, sections={k: round(v, 3) for k, v in tm.items()},
                compliance=[float(v) for v in res["compliance"]])


def cpu_sample_text(r, what):
    return (f"{what}: {r['steps_timed']} timed step(s) after {r['warmup']} warm-up, "
            f"{r['s_per_step']:.2f} s/step measured on this mesh (no extrapolation); oracle port "
            f"with assembly / Jacobi-PCG (scipy cg semantics, rtol 1e-8, x0 = 0, "
            f"{r['cg_iters'][-1]} iterations in the last solve) / element energies in C + OpenMP "
            f"on {r['threads']} threads, Helmholtz filter by scipy cg, rest NumPy")


def same_config_c1_cpu():
    r = cpu_port_run(C1_MESH_SIZE, "oc", 50, C1_WARMUP, C1_STEPS, 0.8)
    return {"cpu_ms_per_step": 1e3 * r["s_per_step"], "cpu_steps_timed": r["steps_timed"],
            "cpu_threads": r["threads"], "cpu_cg_iters": r["cg_iters"],
            "cpu_compliance": r["compliance"]}


def bench_config(mesh_size, n_elem=None, n_dof=None):
    """The `config` object, identical in both arms."""
    return {"workload": workload_name(mesh_size), "mesh_size": mesh_size,
            "method": "LogMOC, vol_frac 0.3, Helmholtz filter, 200-iteration schedules, "
                      "rtol 1e-8, fp64",
            "same_config_c1": "C1: toy_base(0.155) = 52,728 hex, OC defaults, 50-iteration "
                              "schedules; %d timed steps after %d warm-up in both arms"
                              % (C1_STEPS, C1_WARMUP)}


def run_reference(args):
    """--impl reference: the reference path on the host cores.  The reference
    itself cannot be imported (scikit-fem / pyamg are not installed, no network:
    no baseline/_ref), so this is the oracle port (`kind: port`) -- on the SAME
    mesh as the GPU arm, for as many steps as fit the time budget (at least one;
    `steps` reports the steps actually timed, `steps_requested` the flag)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_timed = max(1, min(int(args.steps), CPU_C2_MAX_STEPS))
    r = cpu_port_run(args.mesh_size, "logmoc", 200, 0, n_timed, 0.3, budget_s=CPU_C2_BUDGET_S)
    value = 1.0 / r["s_per_step"]
    sample = cpu_sample_text(r, workload_name(args.mesh_size))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps_timed"], "warmup": r["warmup"],
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1e3 * r["s_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.mesh_size),
        "details": {"n_elem": r["n_elem"], "cg_iters_per_step": r["cg_iters"],
                    "setup_s": r["setup_s"], "sections_s": r["sections"],
                    "compliance": r["compliance"], "host_cores": os.cpu_count()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["threads"], "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_c1:
        line["same_config_c1"] = same_config_c1_cpu()
    emit(line)


# ---------------------------------------------------------------- GPU arm --
def _make_optimizer(sktopt, mesh_size, method="logmoc", max_iters=200):
    tsk = sktopt.mesh.toy_problem.toy_base(mesh_size)
    tmp = tempfile.mkdtemp(prefix="sktopt_bench_")
    if method == "logmoc":
        cfg = sktopt.core.LogMOC_Config(
            dst_path=tmp, max_iters=max_iters, record_times=max(1, max_iters // 10),
            vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3),
            solver_option="cg_pyamg")
        opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
    else:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=max_iters, record_times=max_iters,
                                    solver_option="cg_pyamg")
        opt = sktopt.core.OC_Optimizer(cfg, tsk)
    opt.parameterize()
    opt.export_enabled = False
    return opt


def _timed_steps(torch, opt, steps, barrier):
    """(seconds by CUDA events on the launch stream, wall seconds) of `steps`
    optimiser iterations, barrier + synchronize on both sides."""
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        opt.optimize_steps(1)
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - w0


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import sktopt
    from sktopt._b200 import device as dev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    opt = _make_optimizer(sktopt, args.mesh_size)
    eng = opt.fem.engine
    n_elem, n_dof, nnz = eng.n_elem, eng.n_dof, eng.nnz

    # pinned host buffers of the per-step API edge
    rho_h = torch.empty(n_elem, dtype=torch.float64).pin_memory()
    out_h = torch.empty(n_elem, dtype=torch.float64).pin_memory()

    if args.warmup > 0:
        opt.optimize_steps(args.warmup)
    opt._ensure_state_initialized()
    st = opt._state
    rho_h.copy_(st.rho)
    eng.pcg.set_profile(2)
    n_solves0 = len(eng.pcg_log)
    sampler = ClockSampler(local)

    if args.profile_step:
        # one steady-state step inside a cudaProfilerStart/Stop window:
        #   ncu --profile-from-start off --metrics gpu__time_duration.sum ... bench.py --profile-step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        opt.optimize_steps(1)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---- timed region 1 (`value`): K steps, state resident in HBM, bracketed by
    # barrier + synchronize, timed on the device with CUDA events recorded on the
    # stream every kernel of the path is launched on (torch's current stream)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    t_step, t_wall = _timed_steps(torch, opt, args.steps, barrier)
    launches = dev.launch_count() - launches0
    spmv_ms_sum, spmv_n = eng.pcg.get_profile()
    pcg_iters = [l[0] for l in eng.pcg_log[n_solves0:]]
    eng.pcg.set_profile(0)

    # ---- timed region 2 (`e2e`): the same K steps through the public API with
    # HOST buffers: H2D of rho from pinned memory, one optimiser step, D2H of the
    # new rho and the compliance, every step inside the timed region
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    rho_h.copy_(st.rho)
    comp_last = None
    n_solves_e2e = len(eng.pcg_log)
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        st.rho.copy_(rho_h, non_blocking=True)          # H2D of the step's input
        opt.optimize_steps(1)
        out_h.copy_(st.rho, non_blocking=True)          # D2H of the step's result
        torch.cuda.synchronize()
        comp_last = float(st.compliance)
        # the result buffer is the next step's input buffer (no host-side copy: a
        # multi-threaded torch CPU copy leaves its OpenMP workers spinning on the
        # cores the solver's host threads need, measured +2 ms on the next step)
        rho_h, out_h = out_h, rho_h
    ev1.record()
    barrier()
    t_e2e_wall = time.perf_counter() - w0
    t_e2e = max(ev0.elapsed_time(ev1) * 1e-3, t_e2e_wall)
    pcg_iters_e2e = [l[0] for l in eng.pcg_log[n_solves_e2e:]]
    clocks = sampler.stop() if rank == 0 else None

    t_step, t_e2e, t_wall = max_over_ranks([t_step, t_e2e, t_wall])
    # N > 1: the SAME workload, z-slab-sharded over the N GPUs (strong scaling)
    value = args.steps / t_step
    e2e = args.steps / t_e2e

    # ---- late stage (the hard regime): the same run continued to iteration
    # `--late-iter` of its 200-iteration schedule (p = 3, beta = 2, a 0/1
    # topology with modulus contrast 1e3), then K more timed steps
    late = None
    if args.late_iter > 0:
        done = int(getattr(opt, "_iter_next", 1)) - 1
        if args.late_iter > done:
            opt.optimize_steps(args.late_iter - done)
        n0 = len(eng.pcg_log)
        t_l, _ = _timed_steps(torch, opt, args.late_steps, barrier)
        (t_l,) = max_over_ranks([t_l])
        its = [l[0] for l in eng.pcg_log[n0:]]
        rp = st.rho_projected
        late = {"first_iteration": args.late_iter + 1, "steps": args.late_steps,
                "ms_per_step": 1e3 * t_l / args.late_steps,
                "iters_per_s": args.late_steps / t_l, "pcg_iters_per_step": its,
                "pcg_converged": bool(all(l[1] for l in eng.pcg_log[n0:])),
                "p": float(opt.schedulers.values_as_list(
                    args.late_iter + 1, ["p"], export_log=False, precision=6)[0]),
                "rho_projected_below_0.1": float((rp < 0.1).double().mean()),
                "rho_projected_above_0.9": float((rp > 0.9).double().mean()),
                "compliance": float(st.compliance)}

    # ---- assembled-operator SpMV leg (north_star's "PCG SpMV HBM GB/s vs
    # peak"): K(rho) of the SAME mesh and the current density assembled by the
    # gather kernel, then the node-block TMA SpMV (the PCG's kernel on meshes
    # that are not tensor grids, and on multigrid level 1) timed launch by launch
    # with CUDA events.  One launch streams 2.1 GB >> the 126 MB L2.
    spmv_leg = None
    if world == 1 and eng.dpn == 3 and not args.no_spmv_leg:
        eng.assemble(enforce=True)
        xs = torch.randn(n_dof, dtype=torch.float64, device="cuda")
        ys = torch.empty_like(xs)
        reps = 20
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(reps)]
        for _ in range(3):
            dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, xs, eng.max_deg, out=ys)
        torch.cuda.synchronize()
        for a, b in evs:
            a.record()
            dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, xs, eng.max_deg, out=ys)
            b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) for a, b in evs)
        spmv_leg = {"mean_ms": float(np.mean(ts)), "median_ms": ts[len(ts) // 2], "reps": reps}
        del xs, ys
        eng._pattern = None                      # free the 3 GB assembled operator again
        torch.cuda.empty_cache()

    details = {
        "n_elem": n_elem, "n_dof": n_dof, "nnz": nnz,
        "solver": ("device PCG rtol 1e-8, start vector = Galerkin projection on the last "
                   "%d solutions, operator: " % eng.start_hist
                   + ("matrix-free grid stencil" if eng.matrix_free else "assembled node-block CSR")
                   + ", preconditioner: "
                   + ("geometric multigrid V-cycle (Galerkin coarse operators, damped Jacobi "
                      "sweeps per level %s, exact dense coarsest solve, fp32 level-0 products)"
                      % ",".join(str(v) for v in eng.mg.sweeps)
                      if eng.precond == "mg" else "Jacobi")),
        "pcg_iters_per_step": pcg_iters,
        # the e2e region is the NEXT K optimiser iterations of the same run (it cannot
        # repeat the timed ones: the state has moved on); later iterations need fewer
        # PCG iterations, which is why e2e can exceed `value` despite its copies
        "pcg_iters_per_step_e2e": pcg_iters_e2e,
        "l2": ("inputs larger than L2: one step streams the level-1 operator (0.27 GB) ~4x per PCG "
               "iteration plus ~20 work vectors of 25 MB against the 126 MB L2; nothing is "
               "flushed explicitly"),
        "filter": ("Helmholtz: direct fast-diagonalisation solve (adjoint) + matrix-free PCG "
                   "(forward, fixed nodes)" if world == 1 else
                   "Helmholtz: z-slab-sharded matrix-free PCG, filtered field all-gathered"),
        "parallelism": "single GPU" if world == 1 else (
            f"z-slab sharding over {world} GPUs: matrix-free level 0, multigrid levels "
            f"{[l for l, sh in enumerate(eng.mg.shard) if sh is not None] if eng.mg else []} "
            f"and the Helmholtz filter PCG sharded (NCCL plane exchange + dot all-reduces); "
            f"element-wise stages replicated"),
        "halo": (None if world == 1 else
                 "peer memory (NVLink P2P pulls through cudaIpc, one kernel per exchange)"
                 if getattr(eng.comm, "p2p", False) else "ncclSend / ncclRecv"),
        "rows_per_rank": int(eng.n_local), "halo_dofs": int(getattr(eng, "halo_dofs", 0)),
        "last_compliance": comp_last,
    }

    # ---- same-config pair on C1 (52,728 hex, OC): a configuration BOTH arms
    # complete, fully measured on both sides
    c1 = None
    if world == 1 and not args.no_c1:
        o1 = _make_optimizer(sktopt, C1_MESH_SIZE, "oc", 50)
        o1.optimize_steps(C1_WARMUP)
        t1, _ = _timed_steps(torch, o1, C1_STEPS, barrier)
        c1 = {"gpu_ms_per_step": 1e3 * t1 / C1_STEPS, "gpu_steps_timed": C1_STEPS,
              "gpu_compliance": [float(v) for v in
                                 np.asarray(o1.recorder.as_object().compliance)],
              "gpu_bisection_steps": list(o1.bisection_steps),
              "gpu_pcg_iters": [l[0] for l in o1.fem.engine.pcg_log]}
        del o1

    # ---- C5 (8.08M hex) under N > 1: the configuration the multi-GPU target is
    # quoted on, a few steps next to the strong-scaling headline
    c5 = None
    if world > 1 and args.c5_steps > 0:
        del opt, st
        torch.cuda.empty_cache()
        o5 = _make_optimizer(sktopt, C5_MESH_SIZE)
        o5.optimize_steps(2)
        n5 = len(o5.fem.engine.pcg_log)
        t5, _ = _timed_steps(torch, o5, args.c5_steps, barrier)
        (t5,) = max_over_ranks([t5])
        e5 = o5.fem.engine
        c5 = {"workload": workload_name(C5_MESH_SIZE), "n_elem": e5.n_elem, "n_dof": e5.n_dof,
              "steps": args.c5_steps, "ms_per_step": 1e3 * t5 / args.c5_steps,
              "iters_per_s": args.c5_steps / t5,
              "pcg_iters_per_step": [l[0] for l in e5.pcg_log[n5:]],
              "sharded_mg_levels": [l for l, sh in enumerate(e5.mg.shard) if sh is not None],
              "compliance": float(o5._state.compliance)}

    if rank == 0:
        peak, peak_src = measured_peak()
        spmv_ms = spmv_ms_sum / max(spmv_n, 1)
        n_loc_nodes = int(eng.n_local) // 3
        if eng.matrix_free:
            # dominant kernel: matrix-free K(rho) p (3 launches per PCG iteration:
            # q = A p and the two level-0 products of the V-cycle).  It is bound
            # by the FP64 FMA pipe, not by HBM: 576 DFMA per node row block + 24
            # for the modulus scaling (DESIGN.md "grid operator").
            flops = n_loc_nodes * (576 + 24) * 2
            gridop_traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "gridop_traffic.json")) as f:
                    gridop_traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                gridop_traffic = None
            fp64_peak = dev.fp64_peak_tflops()
            fp64_peak_const = dev.fp64_peak_tflops(const_operand=True)
            achieved = flops / (spmv_ms * 1e-3) / 1e12 if spmv_n else None
            alg_bytes = n_loc_nodes * (24 + 24 + 24 + 1) + n_elem * 8
            roofline = {
                "bound": "fp64",
                "kernel": "hexgrid_apply_x2_kernel<double,0,true> (matrix-free q = K(rho) p + p.q, two nodes per thread, coefficients from shared memory; the V-cycle runs two more per PCG iteration in fp32)",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (achieved / fp64_peak) if achieved else None,
                "peak_source": "DFMA-chain probe run by this bench (sktb_fp64_probe, profiles/r2_fp64_probe.md); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "peak_const_operand": fp64_peak_const,
                "frac_of_const_operand_peak": (achieved / fp64_peak_const) if achieved else None,
                "alg_flops_per_launch": flops, "avg_launch_ms": spmv_ms, "samples": spmv_n,
                "traffic": gridop_traffic if world == 1 else None,
                "hbm": {"alg_bytes_per_launch": alg_bytes,
                        "achieved_GBs": alg_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_n else None,
                        "peak_GBs": peak, "peak_source": peak_src},
                "note": "the assembled-operator SpMV this replaces streamed 8.44 B/nnz (2.1 GB per "
                        "launch, 0.44 ms at the HBM roofline); the matrix-free product moves "
                        "~85 MB.  Round 1's kernel took its coefficients as constant-bank operands "
                        "(half-rate DFMA on B200, peak_const_operand); this one reads them from "
                        "shared memory into registers (full-rate DFMA) and is bound by instruction "
                        "issue at 8 warps / SM (255 registers)",
            }
        else:
            # SURVEY.md 8(d), for the rows this rank owns
            spmv_bytes = nnz * 12 + int(eng.n_local) * 12 + n_dof * 8
            achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_n else None
            roofline = {
                "bound": "hbm", "kernel": "spmv_bsr3_tma_kernel<true> (PCG q=Ap + p.q, node-block columns, cp.async.bulk ring)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                "peak_source": peak_src, "alg_bytes_per_launch": spmv_bytes,
                "avg_launch_ms": spmv_ms, "samples": spmv_n, "traffic": None,
            }
        if spmv_leg is not None:
            b_alg = nnz * 12 + n_dof * 12 + n_dof * 8          # SURVEY 8(d), CSR accounting
            b_fmt = nnz * 8 + (nnz // 9) * 4 + (n_dof // 3) * 4 + n_dof * 16
            tr = None
            try:
                with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
                    tr = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                tr = None
            ms = spmv_leg["mean_ms"]
            roofline["spmv_assembled"] = {
                "bound": "hbm",
                "kernel": "spmv_bsr3_tma_kernel<false> on K(rho) of the same mesh (assembled by "
                          "assemble_kernel), timed alone launch by launch, operand 2.1 GB >> L2",
                "achieved": b_alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": b_alg / (ms * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8TBs": b_alg / (ms * 1e-3) / 1e9 / 8000.0,
                "alg_bytes_per_launch": b_alg, "format_bytes_per_launch": b_fmt,
                "achieved_format_GBs": b_fmt / (ms * 1e-3) / 1e9,
                "frac_format": b_fmt / (ms * 1e-3) / 1e9 / peak,
                "avg_launch_ms": ms, "median_launch_ms": spmv_leg["median_ms"],
                "samples": spmv_leg["reps"], "traffic": tr, "peak_source": peak_src,
                "note": "achieved uses SURVEY 8(d)'s 12 B/nnz CSR accounting; the kernel reads one "
                        "int32 column per 3x3 block (8.44 B/nnz = format bytes), so frac > 1 is "
                        "format compression; frac_format is the kernel's real byte rate vs the "
                        "copy roofline; traffic = ncu dram bytes per launch (profiles/spmv_traffic.json)",
            }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_step / args.steps,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
            "timing": "CUDA events on the launch stream around the K steps, barrier + synchronize "
                      "on both sides, max over ranks; wall_ms_per_step is the host clock around the "
                      "same region",
            "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": bench_config(args.mesh_size),
            "details": details,
            "roofline": roofline,
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": n_elem * 8, "d2h_bytes_per_step": n_elem * 8 + 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if late is not None:
            line["late_stage"] = late
        if c5 is not None:
            line["c5"] = c5
        if world == 1 and not args.no_cpu:
            # the true C2 workload on the host cores: ONE full step, measured
            r = cpu_port_run(args.mesh_size, "logmoc", 200, 0, 1, 0.3)
            line["cpu_baseline"] = {
                "value": 1.0 / r["s_per_step"], "unit": UNIT, "cores": r["threads"],
                "kind": "port", "sample": cpu_sample_text(r, workload_name(args.mesh_size)),
                "sections_s": r["sections"], "host_cores": os.cpu_count()}
            if c1 is not None:
                c1.update(same_config_c1_cpu())
                c1["speedup"] = c1["cpu_ms_per_step"] / c1["gpu_ms_per_step"]
                gc, cc = np.asarray(c1["gpu_compliance"]), np.asarray(c1["cpu_compliance"])
                c1["compliance_max_rel_diff"] = float(np.max(np.abs(gc - cc) / np.abs(cc)))
        if c1 is not None:
            line["same_config_c1"] = c1
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_workload(args):
    """--workload c3 | c4: BASELINE configs 3 and 4 at their named sizes (not the
    headline: a separate JSON line with iters/s, PCG iterations and the SpMV rate)."""
    import torch
    import sktopt
    from sktopt._b200 import device as dev
    from scripts import workloads
    torch.cuda.set_device(0)
    t0 = time.perf_counter()
    if args.workload == "c3":
        tsk = workloads.c3_task(sktopt)
        name = ("C3: 500,610 Kuhn tets (55x41x37 cells, jittered), 2 load cases, mean compliance, "
                "OC defaults, 50-iteration schedules")
    else:
        tsk = workloads.c4_task(sktopt)
        name = ("C4: heat conduction 253x253x32 = 2,048,288 hex, thermal compliance, Robin + "
                "virtual Robin, OC defaults, 50-iteration schedules, intorder 2")
    tmp = tempfile.mkdtemp(prefix="sktopt_bench_")
    cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=50, record_times=50,
                                solver_option="cg_pyamg")
    opt = sktopt.core.OC_Optimizer(cfg, tsk)
    opt.parameterize()
    opt.export_enabled = False
    setup_s = time.perf_counter() - t0
    eng = opt.fem.engine
    opt.optimize_steps(args.warmup)
    eng.pcg.set_profile(1)
    n0 = len(eng.pcg_log)
    launches0 = dev.launch_count()
    t, wall = _timed_steps(torch, opt, args.steps, torch.cuda.synchronize)
    launches = dev.launch_count() - launches0
    ms_sum, n_s = eng.pcg.get_profile()
    peak, peak_src = measured_peak()
    nnz = eng.nnz
    if eng.smg is not None:
        op_bytes = eng.n_dof * (27 * 8 + 8 + 8)        # stencil format: 27 values / row + x + y
        kernel = "dia_apply_kernel<0,true> (27-point stencil format, q = A p + p.q)"
    elif eng.dpn == 3:
        op_bytes = nnz * 8 + (nnz // 9) * 4 + eng.n_dof * 16 + (eng.n_dof // 3) * 4
        kernel = "spmv_bsr3_tma_kernel<true> (node-block columns; operand %.0f MB, L2 resident " \
                 "below 126 MB)" % (op_bytes / 1e6)
    else:
        op_bytes = nnz * 12 + eng.n_dof * 20
        kernel = "spmv_kernel (scalar CSR)"
    ms = ms_sum / max(n_s, 1)
    line = {
        "metric": "optimizer iters/sec", "workload": name, "value": args.steps / t,
        "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
        "n_elem": eng.n_elem, "n_dof": eng.n_dof, "nnz": nnz, "setup_s": setup_s,
        "preconditioner": eng.precond if (eng.mg or eng.smg) else "jacobi",
        "pcg_iters_per_solve": [l[0] for l in eng.pcg_log[n0:]],
        "pcg_converged": bool(all(l[1] for l in eng.pcg_log[n0:])),
        "objective": [float(v) for v in np.asarray(
            getattr(opt.recorder.as_object(), opt._objective_history_name(tsk)))][-args.steps:],
        "bisection_steps": list(opt.bisection_steps)[-args.steps:],
        "roofline": {"bound": "hbm", "kernel": kernel, "bytes_per_launch": op_bytes,
                     "avg_launch_ms": ms, "samples": n_s,
                     "achieved": op_bytes / (ms * 1e-3) / 1e9 if n_s else None,
                     "peak": peak, "unit": "GB/s",
                     "frac": op_bytes / (ms * 1e-3) / 1e9 / peak if n_s else None,
                     "peak_source": peak_src},
        "gpu_launches": int(launches), "dtype": "f64", "data": "synthetic",
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh-size", type=float, default=C2_MESH_SIZE)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-c1", action="store_true", help="skip the C1 same-config pair")
    ap.add_argument("--no-spmv-leg", action="store_true")
    ap.add_argument("--late-iter", type=int, default=150,
                    help="continue the run to this optimiser iteration and time --late-steps "
                         "more there (0: skip)")
    ap.add_argument("--late-steps", type=int, default=10)
    ap.add_argument("--c5-steps", type=int, default=5,
                    help="under --gpus N > 1: timed steps of the 8.08M-element C5 run (0: skip)")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one extra step inside cudaProfilerStart/Stop before the timed region")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2: the headline (default); c3 / c4: BASELINE configs 3 / 4 at size")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "c2":
        run_workload(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
