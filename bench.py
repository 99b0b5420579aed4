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
CPU_SAMPLE_MESH_SIZE = 0.2     # toy1_fine: 40x30x20 = 24,000 hex (largest mesh the reference defines)


def workload_name(mesh_size: float) -> str:
    tag = "C2" if abs(mesh_size - C2_MESH_SIZE) < 1e-12 else "custom"
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
CPU_MAX_TIMED_STEPS = 3        # ~20 s each on the sample mesh: bounds the CPU legs
CPU_MAX_WARMUP_STEPS = 1


def cpu_oracle_step_rate(steps: int, warmup: int, mesh_size: float):
    """The oracle port of the reference path (NumPy/SciPy, scipy cg + Jacobi,
    splu Helmholtz filter) running full LogMOC iterations on a bounded sample
    mesh: at most CPU_MAX_WARMUP_STEPS untimed + CPU_MAX_TIMED_STEPS timed
    iterations, whatever --steps / --warmup ask for, so that the run ends within
    a few minutes.  Returns (iters/s on the sample, n_elem_sample, seconds per
    step, timed steps)."""
    from oracle import mesh as omesh, optim
    o = omesh.toy_base(mesh_size)
    pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                       o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    n_warm = max(0, min(int(warmup), CPU_MAX_WARMUP_STEPS))
    n_timed = max(1, min(int(steps), CPU_MAX_TIMED_STEPS))
    marks = []
    optim.run(pr, "logmoc", max_iters=200, iters=n_warm + n_timed, vol_frac=0.3,
              solver="cg_jacobi", rtol=1e-8, cg_maxiter=20000, step_times=marks)
    dt = (marks[-1] - marks[n_warm]) / n_timed
    return 1.0 / dt, int(o["t"].shape[1]), dt, n_timed


def run_reference(args):
    """--impl reference: the oracle port on the host cores (the reference itself
    cannot be imported here: scikit-fem / pyamg are not installed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_c2 = 1011920
    rate, n_s, dt, n_t = cpu_oracle_step_rate(args.steps, args.warmup, args.cpu_mesh_size)
    value = rate * n_s / n_c2
    sample = (f"oracle LogMOC iteration (scipy cg+Jacobi rtol 1e-8, splu Helmholtz filter) on "
              f"toy_base({args.cpu_mesh_size}) = {n_s} hex ({dt:.2f} s/step, mean of {n_t} timed "
              f"steps: the CPU leg is capped at {CPU_MAX_TIMED_STEPS}), scaled linearly in "
              f"element count to {n_c2} hex (optimistic for the CPU: CG iterations also grow "
              f"with mesh size); scipy is single-threaded")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.mesh_size), "n_elem": n_c2,
                   "sample_mesh_size": args.cpu_mesh_size},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- GPU arm --
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

    tsk = sktopt.mesh.toy_problem.toy_base(args.mesh_size)
    tmp = tempfile.mkdtemp(prefix="sktopt_bench_")
    cfg = sktopt.core.LogMOC_Config(
        dst_path=tmp, max_iters=200, record_times=20,
        vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.3),
        solver_option="cg_pyamg",
    )
    opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
    opt.parameterize()
    opt.export_enabled = False
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
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = dev.launch_count()
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        opt.optimize_steps(1)
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - w0
    launches = dev.launch_count() - launches0
    t_step = ev0.elapsed_time(ev1) * 1e-3
    spmv_ms_sum, spmv_n = eng.pcg.get_profile()
    pcg_iters = [l[0] for l in eng.pcg_log[n_solves0:]]
    eng.pcg.set_profile(0)

    # ---- timed region 2 (`e2e`): the same K steps through the public API with
    # HOST buffers: H2D of rho from pinned memory, one optimiser step, D2H of the
    # new rho and the compliance, every step inside the timed region
    rho_h.copy_(st.rho)
    comp_last = None
    barrier()
    w0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        st.rho.copy_(rho_h, non_blocking=True)          # H2D of the step's input
        opt.optimize_steps(1)
        out_h.copy_(st.rho, non_blocking=True)          # D2H of the step's result
        torch.cuda.synchronize()
        comp_last = float(st.compliance)
        rho_h.copy_(out_h)
    ev1.record()
    barrier()
    t_e2e_wall = time.perf_counter() - w0
    t_e2e = max(ev0.elapsed_time(ev1) * 1e-3, t_e2e_wall)
    clocks = sampler.stop() if rank == 0 else None

    times = torch.tensor([t_step, t_e2e, t_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_step, t_e2e, t_wall = float(times[0]), float(times[1]), float(times[2])
    # N > 1: the SAME workload, its elasticity operator row-sharded over the N
    # GPUs (strong scaling); element-wise stages and the filter are replicated
    value = args.steps / t_step
    e2e = args.steps / t_e2e

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
                "kernel": "hexgrid_apply_shfl_kernel<double,0,true> (matrix-free q = K(rho) p + p.q; the V-cycle runs two more per PCG iteration in fp32)",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (achieved / fp64_peak) if achieved else None,
                "peak_source": "DFMA-chain probe run by this bench (sktb_fp64_probe); "
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
                        "~85 MB.  Its DFMAs take the stiffness coefficient as a constant-bank "
                        "operand, which B200 issues at half the FP64 rate (peak_const_operand, "
                        "measured by sktb_fp64_probe_const): that is the ceiling this formulation "
                        "can reach",
            }
        else:
            # SURVEY.md 8(d), for the rows this rank owns
            spmv_bytes = nnz * 12 + int(eng.n_local) * 12 + n_dof * 8
            achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_n else None
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
            if os.path.exists(tpath):
                try:
                    with open(tpath) as f:
                        traffic = json.load(f).get("dram_bytes_per_launch")
                except Exception:
                    traffic = None
            roofline = {
                "bound": "hbm", "kernel": "spmv_bsr3_tma_kernel<true> (PCG q=Ap + p.q, node-block columns, cp.async.bulk ring)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                "peak_source": peak_src, "alg_bytes_per_launch": spmv_bytes,
                "avg_launch_ms": spmv_ms, "samples": spmv_n,
                "traffic": traffic if world == 1 else None,
                "format_bytes_per_launch": nnz * 8 + (nnz // 9) * 4 + (int(eng.n_local) // 3) * 4
                + int(eng.n_local) * 8 + n_dof * 8,
                "note": "achieved uses SURVEY 8(d) CSR bytes (12 B/nnz); the kernel reads one int32 "
                        "column per 3x3 block (8.44 B/nnz), so a value above the copy roofline is "
                        "format compression, see traffic",
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
            "config": {
                "workload": workload_name(args.mesh_size),
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
                "l2": ("inputs larger than L2: one step streams the level-1 operator (0.27 GB) ~4x per PCG "
                       "iteration plus ~20 work vectors of 25 MB against the 126 MB L2; nothing is "
                       "flushed explicitly"),
                "filter": "Helmholtz: direct fast-diagonalisation solve (adjoint) + matrix-free PCG (forward, fixed nodes)",
                "parallelism": "single GPU" if world == 1 else (
                    f"elasticity operator row-sharded over {world} GPUs (NCCL halo exchange + "
                    f"dot all-reduce); filter/element stages replicated"),
                "rows_per_rank": int(eng.n_local), "halo_dofs": int(getattr(eng, "halo_dofs", 0)),
                "last_compliance": comp_last,
            },
            "roofline": roofline,
            "e2e": {"value": e2e, "unit": UNIT,
                    "h2d_bytes_per_step": n_elem * 8, "d2h_bytes_per_step": n_elem * 8 + 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            rate, n_s, dt, _ = cpu_oracle_step_rate(1, 0, args.cpu_mesh_size)
            line["cpu_baseline"] = {
                "value": rate * n_s / n_elem, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": (f"oracle LogMOC iteration on toy_base({args.cpu_mesh_size}) = {n_s} hex "
                           f"({dt:.2f} s/step, scipy cg+Jacobi, splu filter), scaled linearly in "
                           f"element count to {n_elem} hex (optimistic for the CPU)"),
            }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh-size", type=float, default=C2_MESH_SIZE)
    ap.add_argument("--cpu-mesh-size", type=float, default=CPU_SAMPLE_MESH_SIZE)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spmv-leg", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run one extra step inside cudaProfilerStart/Stop before the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
