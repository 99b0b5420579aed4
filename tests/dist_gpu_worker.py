"""torchrun worker for tests/test_gpu_dist.py: the z-slab-sharded elasticity
solve (matrix-free level 0, sharded multigrid levels, sharded Helmholtz filter)
and the optimiser loop on N ranks against the single-GPU path on the same rank.

SKTOPT_DIST_ONE_GPU=1: every rank uses cuda:0 and the collectives go through the
host shared-memory transport (csrc/comm.cuh) -- the same sharded code paths on a
one-GPU box.  Otherwise one GPU per rank over NCCL."""
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
    one_gpu = os.environ.get("SKTOPT_DIST_ONE_GPU", "0") == "1"
    local = 0 if one_gpu else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import sktopt
    from sktopt._b200 import device as dev
    from sktopt.fea._engine import KE_ELASTIC, get_engine

    mesh_size = float(os.environ.get("SKTOPT_DIST_MESH", "0.25"))
    tsk = sktopt.mesh.toy_problem.toy_base(mesh_size)
    tsk.exlude_dirichlet_from_design()
    rho = np.random.default_rng(5).uniform(0.05, 1.0, tsk.mesh.nelements)
    fem = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
    eng = fem.engine
    assert eng.sharded and eng.comm.world == world
    u = np.zeros((tsk.basis.N, 1))
    c = fem.objectives_multi_load(rho, 3.0, u)
    iters_sharded = eng.pcg_log[-1][0]
    n_sharded_levels = 0 if eng.mg is None else sum(s is not None for s in eng.mg.shard)
    # sharded levels whose V-cycle products stream single-precision values
    n_fp32_sharded = 0 if eng.mg is None else sum(
        1 for l in range(1, eng.mg.n_levels)
        if eng.mg.shard[l] is not None and "vals32" in eng.mg.levels[l])
    slab = eng.slab is not None

    # single-GPU engine on the same rank (replicated), same inputs
    eng1 = get_engine(tsk.basis, tsk.dirichlet_dofs, KE_ELASTIC, tsk.nu, shard=False)
    eng1.set_modulus(dev.to_dev(rho), tsk.E, tsk.E * 1e-3, 3.0)
    eng1.prepare()
    f = dev.to_dev(tsk.neumann_linear[0])
    dev.enforce_rhs(f, None, eng1.dir_mask, None, out=eng1.rhs)
    u1 = eng1.solve(eng1.rhs, 0, 1e-8, None).cpu().numpy()
    iters_single = eng1.pcg_log[-1][0]
    c1 = dev.dot(eng1.rhs, eng1.solution(0))
    err_u = float(np.max(np.abs(u[:, 0] - u1)) / np.max(np.abs(u1)))
    err_c = abs(c[0] - c1) / abs(c1)
    # the sharded matrix rows equal the corresponding rows of the full matrix
    eng.assemble(enforce=True)     # (a matrix-free engine never assembled them)
    eng1.assemble(enforce=True)
    lo = int(eng.col_idx.numel())
    full_vals = eng1.vals.cpu().numpy()
    rp_full = eng1.row_ptr.cpu().numpy()
    s, e = rp_full[eng.row0], rp_full[eng.row0 + eng.n_local]
    assert e - s == lo
    rows_equal = bool(np.array_equal(full_vals[s:e], eng.vals.cpu().numpy()))

    # sharded Helmholtz filter (PCG on z-slabs) against the replicated one
    os.environ["SKTOPT_B200_FILTER_SHARD"] = "1"
    os.environ["SKTOPT_B200_HELMHOLTZ_FD"] = "0"      # adjoint system through the sharded PCG too
    F = sktopt.filters.HelmholtzFilterNodal.from_defaults
    dmask = tsk.design_mask
    f_sh = F(tsk.mesh, tsk.elements_volume, 0.4, dmask)
    y_sh, g_sh = f_sh.forward(rho), f_sh.gradient(-rho)
    filter_sharded = f_sh._device().comm is not None
    os.environ["SKTOPT_B200_FILTER_SHARD"] = "0"
    os.environ.pop("SKTOPT_B200_HELMHOLTZ_FD")
    f_re = F(tsk.mesh, tsk.elements_volume, 0.4, dmask)
    y_re, g_re = f_re.forward(rho), f_re.gradient(-rho)
    os.environ["SKTOPT_B200_FILTER_SHARD"] = "1"
    err_f = float(max(np.max(np.abs(y_sh - y_re)), np.max(np.abs(g_sh - g_re))))

    # optimiser loop: identical on every rank and equal to the single-GPU oracle run;
    # every rank passes the SAME dst_path (rank 0 owns the run directory)
    tmp = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    cfg = sktopt.core.OC_Config(dst_path=os.path.join(tmp[0], "run"), max_iters=4,
                                record_times=4, solver_option="cg_pyamg")
    opt = sktopt.core.OC_Optimizer(cfg, sktopt.mesh.toy_problem.toy_test())
    opt.parameterize()
    opt.optimize()
    comp = np.asarray(opt.recorder.as_object().compliance)
    rho_fin = opt._state.rho.clone()
    if one_gpu:
        rho_fin = rho_fin.cpu()          # gloo gathers host tensors
    gathered = [torch.empty_like(rho_fin) for _ in range(world)]
    dist.all_gather(gathered, rho_fin)
    same = all(bool(torch.equal(g, gathered[0])) for g in gathered)
    dist.barrier()
    files_ok = True
    if rank == 0:
        run = os.path.join(tmp[0], "run")
        files_ok = (os.path.exists(os.path.join(run, "histories.npz"))
                    and os.path.exists(os.path.join(run, "data", "000004-rho.npz"))
                    and os.path.exists(os.path.join(run, "mesh_rho", "info_mesh-00000004.vtu")))

    if rank == 0:
        from oracle import mesh as omesh, optim
        o = omesh.toy_base(1.0)
        pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                           o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
        ref = optim.run(pr, "oc", max_iters=4)
        rel = float(np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])))
        drho = float(np.max(np.abs(rho_fin.cpu().numpy() - ref["rho_final"])))
        p2p = bool(getattr(eng.comm, "p2p", False))
        arena = eng.comm.arena_status() if p2p else (0, 0, 0)
        print(f"DIST_P2P p2p={p2p} arena_bytes={arena[0]} exchanges={arena[1]} err={arena[2]}")
        assert arena[2] == 0
        print(f"DIST_RESULT world={world} slab={slab} sharded_mg_levels={n_sharded_levels} "
              f"fp32_sharded_levels={n_fp32_sharded} "
              f"err_u={err_u:.3e} err_c={err_c:.3e} rows_equal={rows_equal} "
              f"iters={iters_sharded}/{iters_single} filter_sharded={filter_sharded} "
              f"err_filter={err_f:.3e} same_rho={same} loop_rel={rel:.3e} "
              f"loop_drho={drho:.3e} files_ok={files_ok}")
        ok = (err_u <= 1e-6 and err_c <= 1e-8 and rows_equal and same and err_f <= 1e-9
              and abs(iters_sharded - iters_single) <= 2
              and rel <= 1e-6 and drho <= 1e-4 and files_ok)
        print("DIST_OK" if ok else "DIST_FAIL")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
