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


# ---------------------------------------------------------------- lattice meshes --
def _node_graph(t, n):
    nen = t.shape[0]
    r = np.repeat(t, nen, axis=0).ravel()
    c = np.tile(t, (nen, 1)).ravel()
    G = sp.csr_matrix((np.ones(r.size), (r, c)), shape=(n, n))
    G.sum_duplicates()
    G.sort_indices()
    return G.indptr, G.indices


def test_lattice_detection_and_graph():
    """``detect_lattice`` reads the node counts off the node graph alone (hexahedra:
    27-point stencil, Kuhn tetrahedra: 15 points) and rejects other numberings;
    ``lattice_graph`` is the node graph of the hexahedral lattice."""
    from sktopt._fem import MeshHex, MeshTet
    from sktopt.fea._multigrid import detect_lattice, lattice_graph
    for M, dims in ((MeshTet, (6, 5, 4)), (MeshHex, (6, 5, 4)), (MeshTet, (9, 4, 7)),
                    (MeshHex, (4, 4, 4)), (MeshTet, (4, 4, 2)), (MeshHex, (3, 3, 12))):
        m = M.init_tensor(*[np.linspace(0, 1, d) for d in dims])
        rp, ci = _node_graph(m.t, m.p.shape[1])
        assert detect_lattice(rp, ci, m.p.shape[1]) == dims, (M.__name__, dims)
    m = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in (6, 5, 4)])
    n = m.p.shape[1]
    rp, ci = _node_graph(m.t, n)
    rp2, ci2 = lattice_graph(6, 5, 4)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    assert rp2.dtype == np.int32 and ci2.dtype == np.int32
    perm = np.random.default_rng(0).permutation(n)
    assert detect_lattice(*_node_graph(perm[m.t], n), n) is None
    # this package's tetrahedral box (Kuhn split of the tensor grid) qualifies
    from sktopt.mesh.toy_problem import create_box_tet
    mt = create_box_tet(1.0, 1.0, 1.0, 0.2)
    assert detect_lattice(*_node_graph(mt.t, mt.p.shape[1]), mt.p.shape[1]) == (6, 6, 6)
    # a line of cells (fewer than 3 nodes along y) and a tiny graph do not
    m2 = MeshHex.init_tensor(np.linspace(0, 1, 9), np.linspace(0, 1, 2), np.linspace(0, 1, 5))
    assert detect_lattice(*_node_graph(m2.t, m2.p.shape[1]), m2.p.shape[1]) is None


def test_algebraic_galerkin_gather_rule():
    """NumPy restatement of the gather rule of ``galerkin_bsr3_lattice_kernel``
    (csrc/galerkin_bsr.cu): coarse block (I, J) = sum over the fine nodes i that
    interpolate from I and the blocks (i, j) of their rows with j interpolating from
    J of w_iI w_jJ A_ij, fixed fine rows / columns left out, fixed coarse dofs as
    identity.  It must equal P^T A P built from the same axis tables, on jittered
    Kuhn tetrahedra with odd and even node counts."""
    from sktopt._fem import MeshTet
    from sktopt.fea._multigrid import axis_tables, coarse_index_map, lattice_graph
    cells = (5, 4, 3)
    m = MeshTet.init_tensor(*[np.linspace(0, 1, c + 1) for c in cells])
    p = m.p.copy()
    p += np.random.default_rng(1).uniform(-0.03, 0.03, p.shape)
    from sktopt.mesh.utils import fix_tetrahedron_orientation
    t = fix_tetrahedron_orientation(m.t, p)
    n = p.shape[1]
    rho = np.random.default_rng(2).uniform(0.1, 1.0, t.shape[1])
    K = fem.assemble_stiffness(p, t, rho, 1.0, 1e-3, 3.0, 0.3)
    clamp = np.nonzero(m.p[0] < 1e-9)[0]
    D = np.unique((3 * clamp[:, None] + np.arange(3)).ravel())
    A, _ = fem.enforce(K, np.zeros(3 * n), D)
    A = A.tocsr()
    fmask = np.zeros(3 * n, bool)
    fmask[D] = True
    fnx, fny, fnz = (c + 1 for c in cells)
    tabs = [axis_tables(c) for c in cells]
    cn = [(c + 1) // 2 + 1 for c in cells]
    cnx, cny, cnz = cn
    fm = [coarse_index_map(c) for c in cells]
    Iz, Ix, Iy = np.meshgrid(np.arange(cnz), np.arange(cnx), np.arange(cny), indexing="ij")
    fnode = (fm[1][Iy] + fny * fm[0][Ix] + fny * fnx * fm[2][Iz]).ravel()
    cmask = fmask.reshape(-1, 3)[fnode].ravel()
    rp, ci = _node_graph(t, n)
    crp, cci = lattice_graph(cnx, cny, cnz)

    def w_axis(ax, f, J):
        c0, c1, w0, w1, _, _ = tabs[ax]
        w = 0.0
        if c0[f] == J:
            w += w0[f]
        if c1[f] == J and c1[f] != c0[f]:
            w += w1[f]
        return w

    nc = cnx * cny * cnz
    got = sp.lil_matrix((3 * nc, 3 * nc))
    Ad = A.toarray()
    for I in range(nc):
        Iy_, Ix_, Iz_ = I % cny, (I // cny) % cnx, I // (cny * cnx)
        for J in cci[crp[I]:crp[I + 1]]:
            Jy, Jx, Jz = J % cny, (J // cny) % cnx, J // (cny * cnx)
            acc = np.zeros((3, 3))
            for sz in range(3):
                fz = tabs[2][4][sz, Iz_]
                for sx in range(3):
                    fx = tabs[0][4][sx, Ix_]
                    for sy in range(3):
                        fy = tabs[1][4][sy, Iy_]
                        if min(fz, fx, fy) < 0:
                            continue
                        wI = tabs[2][5][sz, Iz_] * tabs[0][5][sx, Ix_] * tabs[1][5][sy, Iy_]
                        i = fy + fny * (fx + fnx * fz)
                        for j in ci[rp[i]:rp[i + 1]]:
                            jy, jx, jz = j % fny, (j // fny) % fnx, j // (fny * fnx)
                            w = w_axis(2, jz, Jz) * w_axis(0, jx, Jx) * w_axis(1, jy, Jy) * wI
                            if w == 0.0:
                                continue
                            blk = Ad[3 * i:3 * i + 3, 3 * j:3 * j + 3].copy()
                            blk[fmask[3 * i:3 * i + 3], :] = 0.0
                            blk[:, fmask[3 * j:3 * j + 3]] = 0.0
                            acc += w * blk
            for a in range(3):
                for b in range(3):
                    v = acc[a, b]
                    if cmask[3 * I + a] or cmask[3 * J + b]:
                        v = 1.0 if (I == J and a == b) else 0.0
                    got[3 * I + a, 3 * J + b] = v
    P = _prolongation(cells)
    free_f = sp.diags((~fmask).astype(float))
    free_c = sp.diags((~cmask).astype(float))
    ref = (free_c @ P.T @ free_f @ A @ free_f @ P @ free_c + sp.diags(cmask.astype(float))).tocsr()
    assert abs(got.tocsr() - ref).max() <= 1e-13 * abs(ref).max()
    # nothing of P^T A P falls outside the 27-point coarse graph
    full = (P.T @ free_f @ A @ free_f @ P).tocoo()
    G = sp.csr_matrix((np.ones(cci.size), cci, crp), shape=(nc, nc))
    assert np.all(G[full.row // 3, full.col // 3] == 1.0)


def test_brick_element_matrix_block_diagonalises_under_its_reflections():
    """Groundwork for the next grid-operator kernel: in the
    symmetry-adapted basis of the brick's three reflections the 24 x 24 element matrix is
    8 blocks of 3 x 3 (scripts/ke0_symmetry.py), for any edge lengths."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "scripts", "ke0_symmetry.py")
    spec = importlib.util.spec_from_file_location("ke0_symmetry", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for h in ((0.0577, 0.0577, 0.0571), (0.1, 0.25, 0.07), (1.0, 1.0, 1.0)):
        off, orth = mod.off_block_ratio(h)
        assert off <= 1e-13 and orth <= 1e-14, (h, off, orth)
