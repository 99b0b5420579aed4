"""CPU (gloo, world_size 2 and 3) tests of the multi-GPU host logic: node-range
partition, halo lists, and the row-sharded Jacobi-PCG protocol (halo exchange
of the full-length direction vector, all-reduced dots, all-gather of the
solution) that csrc/pcg.cu implements over NCCL."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _system():
    from oracle import fem, mesh as omesh
    o = omesh.toy_base(1.0)
    rho = np.random.default_rng(0).uniform(0.2, 1.0, o["t"].shape[1])
    K = fem.assemble_stiffness(o["p"], o["t"], rho, 210e3, 210.0, 3.0, 0.3)
    K_e, f_e = fem.enforce(K, o["force"], o["dirichlet_dofs"])
    G = fem.structural_pattern(o["t"], o["p"].shape[1], 1)
    return K_e.tocsr(), f_e, G.indptr.astype(np.int64), G.indices.astype(np.int64)


def test_partition_and_halo_are_consistent():
    from sktopt._b200 import dist as bdist
    K_e, f_e, node_ptr, node_col = _system()
    n_nodes = node_ptr.size - 1
    for world in (1, 2, 3, 5):
        cuts = bdist.partition_nodes(node_ptr, world)
        assert cuts[0] == 0 and cuts[-1] == n_nodes and np.all(np.diff(cuts) > 0)
        nnz = np.diff(node_ptr[cuts])
        assert nnz.max() <= 1.35 * nnz.mean() + 27
        halos = [bdist.build_halo(node_ptr, node_col, cuts, r, 3) for r in range(world)]
        for r, (peers, so, si, ro, ri) in enumerate(halos):
            assert r not in peers
            lo, hi = 3 * cuts[r], 3 * cuts[r + 1]
            assert np.all((si >= lo) & (si < hi))          # we only send what we own
            assert np.all((ri < lo) | (ri >= hi))          # we only receive ghosts
            for i, pr in enumerate(peers):
                # what r sends to pr is exactly what pr expects from r, same order
                ppeers, _, _, pro, pri = halos[pr]
                j = int(np.nonzero(ppeers == r)[0][0])
                assert np.array_equal(si[so[i]:so[i + 1]], pri[pro[j]:pro[j + 1]])
            # every ghost column of the owned rows is covered
            rows = slice(lo, hi)
            cols = np.unique(K_e[rows].indices)
            ghosts = cols[(cols < lo) | (cols >= hi)]
            assert np.array_equal(np.sort(ri), ghosts)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from sktopt._b200 import dist as bdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K_e, f_e, node_ptr, node_col = _system()
    n = K_e.shape[0]
    cuts = bdist.partition_nodes(node_ptr, world)
    peers, so, si, ro, ri = bdist.build_halo(node_ptr, node_col, cuts, rank, 3)
    lo, hi = 3 * int(cuts[rank]), 3 * int(cuts[rank + 1])
    A = K_e[lo:hi]                                  # owned rows, global columns
    minv = 1.0 / K_e.diagonal()[lo:hi]
    b = f_e[lo:hi]

    def halo(p):
        reqs, bufs = [], []
        for i, pr in enumerate(peers):
            send = torch.from_numpy(np.ascontiguousarray(p[si[so[i]:so[i + 1]]]))
            recv = torch.empty(int(ro[i + 1] - ro[i]), dtype=torch.float64)
            reqs.append(dist.isend(send, int(pr)))
            reqs.append(dist.irecv(recv, int(pr)))
            bufs.append((i, recv, send))
        for q in reqs:
            q.wait()
        for i, recv, _ in bufs:
            p[ri[ro[i]:ro[i + 1]]] = recv.numpy()

    def allsum(*vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.tolist()

    x = np.zeros(hi - lo)
    r = b.copy()
    z = minv * r
    p = np.zeros(n)
    p[lo:hi] = z
    rz, rr, bb = allsum(r @ z, r @ r, b @ b)
    tol2 = (1e-10 ** 2) * bb
    iters = 0
    while rr > tol2 and iters < 5000:
        halo(p)
        q = A @ p
        (pq,) = allsum(p[lo:hi] @ q)
        alpha = rz / pq
        x += alpha * p[lo:hi]
        r -= alpha * q
        z = minv * r
        rz_new, rr = allsum(r @ z, r @ r)
        p[lo:hi] = z + (rz_new / rz) * p[lo:hi]
        rz = rz_new
        iters += 1
    full = [torch.empty(3 * int(cuts[k + 1] - cuts[k]), dtype=torch.float64) for k in range(world)]
    dist.all_gather(full, torch.from_numpy(x)) if len({t.numel() for t in full}) == 1 else None
    if len({t.numel() for t in full}) != 1:
        # variable sizes: gather through broadcasts, like comm_allgatherv
        for k in range(world):
            if k == rank:
                full[k].copy_(torch.from_numpy(x))
            dist.broadcast(full[k], src=k)
    u = torch.cat(full).numpy()
    np.save(os.path.join(out_dir, f"u_{rank}.npy"), u)
    np.save(os.path.join(out_dir, f"it_{rank}.npy"), np.array([iters]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_pcg_protocol_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    K_e, f_e, _, _ = _system()
    u_ref = spla.spsolve(K_e.tocsc(), f_e)
    us = [np.load(tmp_path / f"u_{r}.npy") for r in range(world)]
    its = [int(np.load(tmp_path / f"it_{r}.npy")[0]) for r in range(world)]
    assert len(set(its)) == 1                       # every rank took the same decisions
    for u in us:
        assert np.array_equal(u, us[0])             # replicated result is bit-identical
        assert np.max(np.abs(u - u_ref)) <= 1e-7 * np.abs(u_ref).max()
