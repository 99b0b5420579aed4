"""torchrun worker for tests/test_gpu_dist.py: row-sharded elasticity solve and
optimiser loop on N GPUs against the single-GPU path on the same rank."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import sktopt
    from sktopt._b200 import device as dev
    from sktopt.fea._engine import KE_ELASTIC, get_engine

    tsk = sktopt.mesh.toy_problem.toy_base(0.45)
    tsk.exlude_dirichlet_from_design()
    rho = np.random.default_rng(5).uniform(0.05, 1.0, tsk.mesh.nelements)
    fem = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
    assert fem.engine.sharded and fem.engine.comm.world == world
    u = np.zeros((tsk.basis.N, 1))
    c = fem.objectives_multi_load(rho, 3.0, u)
    iters_sharded = fem.engine.pcg_log[-1][0]

    # single-GPU engine on the same rank (replicated), same inputs
    eng1 = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu, shard=False)
    eng1.set_modulus(dev.to_dev(rho), tsk.E, tsk.E * 1e-3, 3.0)
    eng1.assemble(enforce=True)
    eng1.update_preconditioner()
    f = dev.to_dev(tsk.neumann_linear[0])
    dev.enforce_rhs(f, None, eng1.dir_mask, None, out=eng1.rhs)
    u1 = eng1.solve(eng1.rhs, 0, 1e-8, None).cpu().numpy()
    c1 = dev.dot(eng1.rhs, eng1.solution(0))
    err_u = float(np.max(np.abs(u[:, 0] - u1)) / np.max(np.abs(u1)))
    err_c = abs(c[0] - c1) / abs(c1)
    # the sharded matrix rows equal the corresponding rows of the full matrix
    eng = fem.engine
    eng.assemble(enforce=True)     # (a matrix-free engine never assembled them)
    lo = int(eng.col_idx.numel())
    full_vals = eng1.vals.cpu().numpy()
    rp_full = eng1.row_ptr.cpu().numpy()
    s, e = rp_full[eng.row0], rp_full[eng.row0 + eng.n_local]
    assert e - s == lo
    rows_equal = bool(np.array_equal(full_vals[s:e], eng.vals.cpu().numpy()))

    # optimiser loop: identical on every rank and equal to the single-GPU oracle run
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=os.path.join(tmp, f"r{rank}"), max_iters=4,
                                    record_times=4, solver_option="cg_pyamg")
        opt = sktopt.core.OC_Optimizer(cfg, sktopt.mesh.toy_problem.toy_test())
        opt.parameterize()
        opt.optimize()
        comp = np.asarray(opt.recorder.as_object().compliance)
        rho_fin = opt._state.rho.clone()
    gathered = [torch.empty_like(rho_fin) for _ in range(world)]
    dist.all_gather(gathered, rho_fin)
    same = all(bool(torch.equal(g, gathered[0])) for g in gathered)

    if rank == 0:
        from oracle import mesh as omesh, optim
        o = omesh.toy_base(1.0)
        pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                           o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
        ref = optim.run(pr, "oc", max_iters=4)
        rel = float(np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])))
        drho = float(np.max(np.abs(rho_fin.cpu().numpy() - ref["rho_final"])))
        print(f"DIST_RESULT world={world} err_u={err_u:.3e} err_c={err_c:.3e} "
              f"rows_equal={rows_equal} iters={iters_sharded} same_rho={same} "
              f"loop_rel={rel:.3e} loop_drho={drho:.3e}")
        ok = (err_u <= 1e-6 and err_c <= 1e-8 and rows_equal and same
              and rel <= 1e-6 and drho <= 1e-4)
        print("DIST_OK" if ok else "DIST_FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
