"""Generates tests/golden/c1_oc50_oracle.npz: BASELINE config 1 at its TRUE size
(toy_base(0.155) = 52 x 39 x 26 = 52,728 hex, 171,720 DOF, OC defaults, 50
iterations) run through the CPU oracle.  Takes tens of minutes on one core, so
it is run offline and the result is committed; tests/test_gpu_c1.py compares the
CUDA path with it at the north-star tolerances (compliance history <= 1e-6
relative, densities after 50 iterations <= 1e-4 L-inf).

The oracle solves K u = f with scipy cg + Jacobi at rtol 1e-11 (three decades
tighter than the product's 1e-8) so that the fixture carries no solver noise of
its own.  Stored: compliance[50], vol_error[50], bisection_steps[50], rho_final
(float64) and every 10th density field as float32 (for locating a divergence).

    python tests/golden/make_c1_fixture.py [--iters 50] [--mesh-size 0.155]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import mesh as omesh, optim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--mesh-size", type=float, default=0.155)
    ap.add_argument("--rtol", type=float, default=1e-11)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c1_oc50_oracle.npz"))
    a = ap.parse_args()
    o = omesh.toy_base(a.mesh_size)
    pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                       o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    tm, marks = {}, []
    t0 = time.perf_counter()
    ref = optim.run(pr, "oc", max_iters=50, iters=a.iters, solver="cg_jacobi", rtol=a.rtol,
                    cg_maxiter=200000, timings=tm, step_times=marks)
    wall = time.perf_counter() - t0
    print("wall %.1f s" % wall, {k: round(v, 1) for k, v in tm.items()})
    rho_hist = np.stack(ref["rho"][9::10]).astype(np.float32) if a.iters >= 10 else np.zeros((0, 0), np.float32)
    np.savez_compressed(
        a.out, mesh_size=a.mesh_size, n_elem=o["t"].shape[1],
        compliance=np.asarray(ref["compliance"]), vol_error=np.asarray(ref["vol_error"]),
        bisection_steps=np.asarray(ref["bisection_steps"]), cg_iters=np.asarray(ref["cg_iters"]),
        rho_final=ref["rho_final"], rho_every10_design=rho_hist,
        step_seconds=np.diff(np.asarray(marks)), oracle_rtol=a.rtol)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
