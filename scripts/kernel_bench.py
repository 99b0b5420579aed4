"""Per-kernel timing of the hot path on one GPU (CUDA events, L2 flushed or
inputs > L2).  Usage: python scripts/kernel_bench.py [mesh_size] [reps]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
import sktopt  # noqa: E402
from sktopt._b200 import device as dev  # noqa: E402
from sktopt.fea._engine import KE_ELASTIC, get_engine  # noqa: E402


def timeit(fn, reps, flush=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(reps)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for a, b in ev:
        if flush is not None:
            dev.flush_l2(flush)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    h = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0577
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    t0 = time.time()
    tsk = sktopt.mesh.toy_problem.toy_base(h)
    tsk.exlude_dirichlet_from_design()
    t_host = time.time() - t0
    t0 = time.time()
    eng = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
    torch.cuda.synchronize()
    t_dev = time.time() - t0
    n, ne = eng.n_dof, eng.n_elem
    nnz = eng.vals.numel()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=dev.F64, device="cuda")
    rho = dev.to_dev(np.random.default_rng(0).uniform(0.2, 1.0, ne))
    out = {"mesh_size": h, "n_elem": ne, "n_dof": n, "nnz": nnz,
           "host_setup_s": t_host, "device_setup_s": t_dev,
           "n_class": eng.dm.n_class}

    eng.set_modulus(rho, tsk.E, tsk.E * 1e-3, 3.0)
    ms, best = timeit(lambda: eng.assemble(enforce=True), reps, flush)
    asm_bytes = nnz * 8 + ne * (8 * 4 + 8) + eng.dm.n_nodes * 24
    out["assemble"] = {"ms": ms, "best_ms": best, "alg_GBps": asm_bytes / ms / 1e6}
    eng.update_preconditioner()

    x = torch.randn(n, dtype=dev.F64, device="cuda")
    y = torch.empty_like(x)
    ms, best = timeit(lambda: eng.spmv(x, out=y), reps, flush)
    spmv_bytes = nnz * 12 + n * 12 + n * 8
    out["spmv"] = {"ms": ms, "best_ms": best, "alg_bytes": spmv_bytes,
                   "alg_GBps": spmv_bytes / ms / 1e6, "best_GBps": spmv_bytes / best / 1e6}

    ms, best = timeit(lambda: dev.spmv_bsr3(eng.node_ptr_loc, eng.node_col_loc, eng.vals, x, out=y),
                      reps, flush)
    bsr_bytes = nnz * 8 + (nnz // 9) * 4 + (n // 3) * 4 + n * 16
    out["spmv_bsr3"] = {"ms": ms, "best_ms": best, "alg_GBps_csr_accounting": spmv_bytes / ms / 1e6,
                        "format_bytes": bsr_bytes, "format_GBps": bsr_bytes / ms / 1e6}

    ms, best = timeit(lambda: dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, x,
                                                eng.max_deg, out=y), reps, flush)
    out["spmv_bsr3_tma"] = {"ms": ms, "best_ms": best, "alg_GBps_csr_accounting": spmv_bytes / ms / 1e6,
                            "format_GBps": bsr_bytes / ms / 1e6}

    u = torch.randn(n, dtype=dev.F64, device="cuda")
    e = torch.empty(ne, dtype=dev.F64, device="cuda")
    ms, best = timeit(lambda: eng.energy(u, out=e), reps, flush)
    en_bytes = ne * (8 * 4 + 16) + n * 8
    out["energy"] = {"ms": ms, "best_ms": best, "alg_GBps": en_bytes / ms / 1e6}

    # PCG: fixed number of iterations (rtol=0 never converges)
    f = dev.to_dev(tsk.neumann_linear[0])
    dev.enforce_rhs(f, None, eng.dir_mask, None, out=eng.rhs)
    xs = eng.solution(0)
    iters = 200
    eng.warm_start = False
    torch.cuda.synchronize()
    t0 = time.time()
    eng.pcg.solve(eng.node_ptr_loc, eng.node_col_loc, eng.vals, eng.inv_diag, eng.rhs, xs,
                  dpn_hint=3, rtol=0.0, maxiter=iters, use_x0=False, check_every=50, block3=True, max_deg=eng.max_deg)
    torch.cuda.synchronize()
    dt = time.time() - t0
    pcg_bytes = spmv_bytes + 16 * n * 8
    out["pcg_iter"] = {"ms": dt / iters * 1e3, "iters": eng.pcg.last_iters,
                       "alg_GBps": pcg_bytes / (dt / iters) / 1e9}
    # real solve to rtol 1e-8
    t0 = time.time()
    eng.pcg.solve(eng.node_ptr_loc, eng.node_col_loc, eng.vals, eng.inv_diag, eng.rhs, xs,
                  dpn_hint=3, rtol=1e-8, maxiter=60000, use_x0=False, check_every=50, block3=True, max_deg=eng.max_deg)
    torch.cuda.synchronize()
    out["pcg_solve"] = {"s": time.time() - t0, "iters": eng.pcg.last_iters,
                        "converged": eng.pcg.last_converged, "relres": eng.pcg.last_relres}

    # matrix-free operator, multigrid and the uniform energy kernel (tensor grids)
    eng2 = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu)
    if eng2.matrix_free:
        eng2.set_modulus(rho, tsk.E, tsk.E * 1e-3, 3.0)
        eng2.prepare()
        n_nodes = eng2.n_dof // 3
        ms, best = timeit(lambda: eng2.spmv(x, out=y), reps, flush)
        out["gridop_fp64"] = {"ms": ms, "best_ms": best,
                              "TFLOPs": n_nodes * 1200 / ms / 1e9}
        if eng2.mg is not None:
            r = torch.randn(n, dtype=dev.F64, device="cuda")
            z = torch.empty_like(r)
            ms, best = timeit(lambda: eng2.mg.vcycle(r, z), reps, flush)
            out["mg_vcycle"] = {"ms": ms, "best_ms": best, "levels": eng2.mg.n_levels,
                                "sweeps": eng2.mg.sweeps}
            ms, best = timeit(lambda: eng2.mg.setup(), max(3, reps // 4), flush)
            out["mg_setup"] = {"ms": ms, "best_ms": best}
        ms, best = timeit(lambda: eng2.energy(u, out=e), reps, flush)
        out["energy_uniform"] = {"ms": ms, "best_ms": best, "alg_GBps": en_bytes / ms / 1e6}

    # Helmholtz filter
    filt = sktopt.filters.HelmholtzFilterNodal.from_defaults(
        tsk.mesh, tsk.elements_volume, 0.01, design_mask=tsk.design_mask)
    r = dev.to_dev(np.full(ne, 0.5))
    filt.forward(r)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(5):
        filt.forward(r)
    torch.cuda.synchronize()
    st = filt._dev_state
    out["helmholtz_forward"] = {"ms": (time.time() - t0) / 5 * 1e3,
                                "pcg_iters": st.solve_iters[-5:]}
    if st.grid is not None:
        xn = torch.randn(st.n_nodes, dtype=dev.F64, device="cuda")
        yn = torch.empty_like(xn)
        ms, best = timeit(lambda: st.gop_A.apply(xn, out=yn), reps, flush)
        out["scalar_stencil"] = {"ms": ms, "best_ms": best,
                                 "alg_GBps": st.n_nodes * 17 / ms / 1e6}
        if st.fd is not None:
            ms, best = timeit(lambda: st.fd.solve(xn, out=yn), reps, flush)
            out["helmholtz_direct_solve"] = {"ms": ms, "best_ms": best}
        g = filt.gradient(-r)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            filt.gradient(-r)
        torch.cuda.synchronize()
        out["helmholtz_gradient"] = {"ms": (time.time() - t0) / 5 * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
