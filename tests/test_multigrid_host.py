"""CPU tests of the host-side tables the multigrid kernels consume
(sktopt/fea/_multigrid.py): 1-D transfer tables, parent/child tables and the
trilinear child->parent weights.  The element-wise Galerkin rule the device
set-up implements, Ke_E = sum_c Q_c^T Ke_c Q_c, assembled over the coarse grid,
must equal P^T K P with the prolongation built from the same axis tables -- on a
grid with odd and even cell counts (half-width parents)."""
import numpy as np
import scipy.sparse as sp

from oracle import fem


def _assemble(t, n_nodes, Ke):
    """Scatter-add of per-element 24x24 matrices (ne, 24, 24)."""
    ne = t.shape[1]
    dofs = (3 * t.T.astype(np.int64)[:, :, None] + np.arange(3)[None, None, :]).reshape(ne, 24)
    rows = np.repeat(dofs, 24, axis=1).ravel()
    cols = np.tile(dofs, (1, 24)).ravel()
    return sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(3 * n_nodes, 3 * n_nodes)).tocsr()


def _prolongation(fine_cells):
    from sktopt.fea._multigrid import axis_tables
    mats = []
    for n in fine_cells:
        c0, c1, w0, w1, fT, wT = axis_tables(n)
        nc = (n + 1) // 2
        P = np.zeros((n + 1, nc + 1))
        for i in range(n + 1):
            P[i, c0[i]] += w0[i]
            P[i, c1[i]] += w1[i]
        assert np.allclose(P.sum(axis=1), 1.0)                      # partition of unity
        # the transposed tables list the same non-zeros, column by column
        PT = np.zeros_like(P)
        for c in range(nc + 1):
            for s in range(3):
                if fT[s, c] >= 0:
                    PT[fT[s, c], c] += wT[s, c]
        assert np.array_equal(P, PT)
        mats.append(sp.csr_matrix(P))
    Px, Py, Pz = mats
    # node = iy + npy (ix + npx iz)  ->  kron(Pz, kron(Px, Py)), three dofs per node
    return sp.kron(sp.kron(Pz, sp.kron(Px, Py)), sp.eye(3)).tocsr()


def test_elementwise_galerkin_equals_PtKP():
    from sktopt._fem import MeshHex
    from sktopt.fea._multigrid import (child_tables, coarse_index_map, detect_tensor_grid,
                                       q_tables, vertex_bits)
    fine_cells = (5, 4, 3)                                         # odd, even, odd
    axes = [np.linspace(0.0, 0.3 * n, n + 1) for n in fine_cells]
    mesh = MeshHex.init_tensor(*axes)
    assert all(np.array_equal(a, b) for a, b in zip(detect_tensor_grid(mesh), axes))
    bits = vertex_bits(mesh)
    assert sorted((bits @ np.array([1, 2, 4])).tolist()) == list(range(8))
    rng = np.random.default_rng(0)
    E = rng.uniform(0.01, 1.0, mesh.nelements)
    nu = 0.3
    lam, mu = nu / ((1 + nu) * (1 - 2 * nu)), 0.5 / (1 + nu)
    Ke0 = fem.elasticity_ke(mesh.p, mesh.t[:, :1], np.array([lam]), np.array([mu]), 2)[0]
    Kf = _assemble(mesh.t, mesh.nvertices, E[:, None, None] * Ke0[None])
    # coarse grid = every other node (the last one kept): ceil-halving
    caxes = [a[coarse_index_map(a.size - 1)] for a in axes]
    coarse_cells = tuple(a.size - 1 for a in caxes)
    assert coarse_cells == (3, 2, 2)
    cmesh = MeshHex.init_tensor(*caxes)
    assert np.array_equal(vertex_bits(cmesh), bits)
    child, ptype = child_tables(fine_cells, coarse_cells)
    # every fine element has exactly one parent
    kids = child[child >= 0]
    assert np.array_equal(np.sort(kids), np.arange(mesh.nelements))
    Q = q_tables(bits)                                              # [type][child][a][A]
    KeC = np.zeros((cmesh.nelements, 24, 24))
    eye3 = np.eye(3)
    for Ec in range(cmesh.nelements):
        for ch in range(8):
            e = child[ch, Ec]
            if e < 0:
                continue
            Qv = np.kron(Q[ptype[Ec], ch], eye3)                   # (24 child dofs, 24 parent dofs)
            KeC[Ec] += Qv.T @ (E[e] * Ke0) @ Qv
    Kc = _assemble(cmesh.t, cmesh.nvertices, KeC)
    P = _prolongation(fine_cells)
    ref = (P.T @ Kf @ P).toarray()
    assert np.max(np.abs(Kc.toarray() - ref)) <= 1e-12 * np.abs(ref).max()
    # rows of Q: each child vertex is a convex combination of parent vertices
    for ty in range(8):
        for ch in range(8):
            rows = Q[ty, ch].sum(axis=1)
            assert np.allclose(rows[rows > 0], 1.0)


def test_chebyshev_coefficients_reproduce_the_polynomial():
    """d_k = c1 d_{k-1} + c2 r_k applied to a diagonal system equals the Chebyshev
    residual polynomial: after `degree` steps the error is damped on [lmin, lmax]
    by at least the Chebyshev bound."""
    from sktopt.fea._multigrid import chebyshev_coefficients
    lmin, lmax, deg = 0.3, 2.0, 5
    c1, c2 = chebyshev_coefficients(lmax, lmin, deg)
    lam = np.linspace(lmin, lmax, 201)
    b = np.ones_like(lam)
    x = np.zeros_like(lam)
    d = np.zeros_like(lam)
    for k in range(deg):
        r = b - lam * x
        d = (c1[k] * d if k else 0.0) + c2[k] * r
        x = x + d
    err = np.abs(1.0 - lam * x)                                     # residual polynomial
    sigma = (lmax + lmin) / (lmax - lmin)
    bound = 1.0 / np.cosh(deg * np.arccosh(sigma))
    assert err.max() <= bound * (1.0 + 1e-9)


def test_slab_sharding_plan_covers_what_every_rank_needs():
    """plan_slab_sharding (host logic of the sharded multigrid set-up): ownership is
    a partition by whole planes on every level, every element matrix a rank needs
    -- for the rows it assembles, for its share of the gathered level and for the
    Galerkin products of the coarser sharded levels -- lies in its element range."""
    from sktopt._b200.dist import partition_planes
    from sktopt.fea._multigrid import child_tables, coarse_index_map, plan_slab_sharding
    for cells, world in (((20, 14, 33), 2), ((17, 9, 40), 3), ((12, 10, 64), 4), ((8, 6, 97), 8)):
        coords = [tuple(np.linspace(0, 1, n + 1) for n in cells)]
        while max(c.size - 1 for c in coords[-1]) > 2:
            coords.append(tuple(a[coarse_index_map(a.size - 1)] for a in coords[-1]))
        cuts0 = partition_planes(cells[2] + 1, world)
        plans = [plan_slab_sharding(coords, cuts0, world, r, 300) for r in range(world)]
        first_rep = plans[0][1]
        assert first_rep >= 2, (cells, world)              # at least level 1 is sharded here
        for l in range(1, first_rep):
            npl = coords[l][0].size * coords[l][1].size
            n_nodes = npl * coords[l][2].size
            ranges = [(p[0][l]["node0"], p[0][l]["node1"]) for p in plans]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_nodes
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            assert all((b - a) % npl == 0 and (b - a) >= 2 * npl for a, b in ranges)
        for r, (shard, lr, gp, zc) in enumerate(plans):
            assert lr == first_rep and gp["level"] == lr
            assert gp["cuts"][0] == 0 and gp["cuts"][-1] == coords[lr][2].size - 1
            # element planes needed on the gathered level -> children on the level below, ...
            need = set(range(int(gp["cuts"][r]), int(gp["cuts"][r + 1])))
            for l in range(lr - 1, 0, -1):
                cx, cy, cz = (c.size - 1 for c in coords[l])
                pe = cx * cy
                sh = shard[l]
                e_lo, e_hi = sh["elem0"] // pe, sh["elem1"] // pe
                assert sh["elem0"] % pe == 0 and sh["elem1"] % pe == 0
                children = {z for E in need for z in (2 * E, 2 * E + 1) if z < cz}
                z0, z1 = sh["node0"] // sh["plane"], sh["node1"] // sh["plane"]
                touching = {z for z in range(cz) if z + 1 >= z0 and z < z1}   # nodes z, z+1
                assert children | touching <= set(range(e_lo, e_hi)), (cells, world, r, l)
                need = set(range(e_lo, e_hi))
        # child tables agree with the "2E, 2E+1" rule used above
        fine = tuple(c.size - 1 for c in coords[0])
        coarse = tuple(c.size - 1 for c in coords[1])
        child, _ = child_tables(fine, coarse)
        ez_c = np.arange(child.shape[1]) // (coarse[0] * coarse[1])
        for ch in range(8):
            ok = child[ch] >= 0
            ez_f = child[ch][ok] // (fine[0] * fine[1])
            assert np.all((ez_f == 2 * ez_c[ok]) | (ez_f == 2 * ez_c[ok] + 1))
