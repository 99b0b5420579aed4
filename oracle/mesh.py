"""CPU ORACLE (test infrastructure): independent restatement of the reference's
box meshes and toy tasks (mesh/toy_problem.py:10-144, mesh/task_common.py,
mesh/task_elastic.py:15-81) as plain arrays.  PARITY UNPINNED (see fem.py).

Used to cross-check the product's host-side task construction: node / DOF
numbering, connectivity, Dirichlet DOFs, design set and load vectors must be
bit-identical between the two independent implementations.
"""
from __future__ import annotations

import numpy as np

from . import fem


def box_hex(x_len, y_len, z_len, h):
    """create_box_hex (mesh/toy_problem.py:10-37) with skfem's
    MeshHex.init_tensor numbering: node = iy + npy*ix +
    npy*npx*iz, element = ey + ny*ex + ny*nx*ez, local vertex order
    (0,0,0) +y +x +z +xy +yz +xz +xyz; then fix_hexahedron_orientation
    (mesh/utils.py:18-54)."""
    nx, ny, nz = (int(np.ceil(L / h)) for L in (x_len, y_len, z_len))
    xs, ys, zs = (np.linspace(0, L, n + 1) for L, n in ((x_len, nx), (y_len, ny), (z_len, nz)))
    npx, npy, npz = nx + 1, ny + 1, nz + 1
    p = np.empty((3, npx * npy * npz))
    for iz in range(npz):
        for ix in range(npx):
            base = npy * ix + npy * npx * iz
            p[0, base:base + npy] = xs[ix]
            p[1, base:base + npy] = ys
            p[2, base:base + npy] = zs[iz]
    ey, ex, ez = np.meshgrid(np.arange(ny), np.arange(nx), np.arange(nz), indexing="ij")
    # element index: ey fastest, then ex, then ez
    order = np.argsort((ey + ny * ex + ny * nx * ez).ravel(), kind="stable")
    ey, ex, ez = ey.ravel()[order], ex.ravel()[order], ez.ravel()[order]
    node = lambda jy, jx, jz: jy + npy * jx + npy * npx * jz
    offs = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    t = np.stack([node(ey + dy, ex + dx, ez + dz) for dy, dx, dz in offs]).astype(np.int32)
    # orientation fix: swap local 1 and 3 where (p1-p0)x(p3-p0).(p4-p0) < 0
    v1 = p[:, t[1]] - p[:, t[0]]
    v2 = p[:, t[3]] - p[:, t[0]]
    v3 = p[:, t[4]] - p[:, t[0]]
    neg = np.einsum("de,de->e", np.cross(v1, v2, axis=0), v3) < 0
    t1 = t[1].copy()
    t[1, neg] = t[3, neg]
    t[3, neg] = t1[neg]
    return p, t


_HEX_FACES = ((0, 1, 4, 2), (3, 5, 7, 6), (0, 1, 5, 3), (2, 4, 7, 6), (0, 2, 6, 3), (1, 4, 7, 5))


def hex_facets(t):
    """All facets (interior too) as unique sorted vertex 4-tuples, with one
    cyclically ordered copy for geometry.  Returns (sorted (4,nf), cyclic (4,nf))."""
    cyc = np.hstack([t[list(f)] for f in _HEX_FACES]).astype(np.int64)
    srt = np.sort(cyc, axis=0)
    _, first = np.unique(srt, axis=1, return_index=True)
    return srt[:, first], cyc[:, first]


def in_box(x, xr, yr, zr):
    """utils.get_points_in_range (mesh/utils.py:8-15): closed intervals."""
    return ((x[0] >= xr[0]) & (x[0] <= xr[1]) & (x[1] >= yr[0]) & (x[1] <= yr[1])
            & (x[2] >= zr[0]) & (x[2] <= zr[1]))


def quad_facet_load(p, cyc, value):
    """asm(LinearForm(value*v), FacetBasis): 2x2 Gauss on each bilinear quad
    facet; returns per-node scalar load (n_nodes,) and the total facet area."""
    g, w = fem.gauss_unit(2)
    F = np.zeros(p.shape[1])
    area = 0.0
    x = p[:, cyc]                                       # (3, 4, nf) cyclic corners
    for r, wr in zip(g, w):
        for s, ws in zip(g, w):
            N = np.array([(1 - r) * (1 - s), r * (1 - s), r * s, (1 - r) * s])
            dNr = np.array([-(1 - s), (1 - s), s, -s])
            dNs = np.array([-(1 - r), -r, r, (1 - r)])
            tr = np.einsum("daf,a->df", x, dNr)
            ts = np.einsum("daf,a->df", x, dNs)
            jac = np.linalg.norm(np.cross(tr, ts, axis=0), axis=0)
            area += float(np.sum(jac) * wr * ws)
            for a in range(4):
                np.add.at(F, cyc[a], value * N[a] * jac * wr * ws)
    return F, area


def toy_base(h):
    """toy_base (mesh/toy_problem.py:51-91) + FEMDomain.from_facets/from_nodes
    (mesh/task_common.py:105-347) + assemble_surface_forces
    (mesh/task_elastic.py:15-81) as a dict of arrays.  The design set already
    has the Dirichlet elements removed (common_density.py:546-547)."""
    x_len, y_len, z_len, eps = 8.0, 6.0, 4.0, 1.2
    p, t = box_hex(x_len, y_len, z_len, h)
    srt, cyc = hex_facets(t)
    mid = p[:, srt].mean(axis=1)
    dir_f = np.nonzero(in_box(mid, (0.0, 0.03), (0.0, y_len), (0.0, z_len)))[0]
    frc_f = np.nonzero(in_box(mid, (x_len - eps, x_len + 0.1),
                              (y_len * 2 / 5, y_len * 3 / 5), (z_len - eps, z_len)))[0]
    cen = p[:, t].mean(axis=1)
    design = np.nonzero(in_box(cen, (0.0, x_len), (0.0, y_len), (0.0, z_len)))[0]
    dir_nodes = np.unique(srt[:, dir_f])
    frc_nodes = np.unique(srt[:, frc_f])
    dir_dofs = np.unique((3 * dir_nodes[:, None] + np.arange(3)[None, :]).ravel())
    touching = lambda nodes: np.nonzero(np.isin(t, nodes).any(axis=0))[0]
    dir_elems, frc_elems = touching(dir_nodes), touching(frc_nodes)
    design = design[~np.isin(design, frc_elems)]          # task_common.py:211-221
    fixed = np.setdiff1d(np.arange(t.shape[1]), design)
    design = design[~np.isin(design, dir_elems)]          # exlude_dirichlet_from_design
    Fz, area = quad_facet_load(p, cyc[:, frc_f], 1.0)
    force = np.zeros(3 * p.shape[1])
    force[2::3] = (-100.0 / area) * Fz                    # 'u^3', value/A traction
    return dict(p=p, t=t, dirichlet_dofs=dir_dofs, dirichlet_nodes=dir_nodes,
                force=force, design=design, fixed=fixed,
                pinned=np.concatenate([dir_elems, frc_elems]),
                volumes=fem.element_volumes(p, t), E=210e3, nu=0.3)
