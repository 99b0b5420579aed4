"""Material interpolation, element volumes and matrix assembly entry points.

Mirrors the public names of reference ``fea/composer.py``.  The assembly
functions run on the GPU: K(rho) = sum_e E_e * Ke0_e where Ke0 is the unit
element matrix (material factors out of the reference's bilinear forms,
``fea/composer.py:80-98`` and ``:136-141``), gathered into CSR through the
precomputed contributor map (``csrc/mesh.cu``).
"""
from __future__ import annotations

from typing import Callable

import numpy as np
import scipy.sparse as sp

from sktopt._b200 import lib as _lib_mod
from sktopt._fem import MeshHex, MeshTet


def simp_interpolation(rho, E0, Emin, p):
    """E = Emin + (E0 - Emin) rho^p  (reference ``fea/composer.py:19-22``)."""
    return Emin + (E0 - Emin) * (rho ** p)


def ramp_interpolation(rho, E0, Emin, p):
    """E = Emin + (E0 - Emin) rho / (1 + p (1 - rho))  (``fea/composer.py:25-39``)."""
    return Emin + (E0 - Emin) * (rho / (1.0 + p * (1.0 - rho)))


simp_interpolation_numba = simp_interpolation
ramp_interpolation_numba = ramp_interpolation


def lam_mu(E, nu):
    return (nu * E) / ((1.0 + nu) * (1.0 - 2.0 * nu)), E / (2.0 * (1.0 + nu))


def is_ramp(elem_func: Callable) -> bool:
    if elem_func in (simp_interpolation, None):
        return False
    if elem_func is ramp_interpolation:
        return True
    raise NotImplementedError(
        "the B200 backend implements the SIMP and RAMP interpolations only"
    )


# ------------------------------------------------------------ element volumes
_HEX_TETS = ((0, 1, 3, 4), (1, 2, 3, 6), (1, 5, 6, 4), (3, 6, 7, 4),
             (1, 3, 6, 4), (1, 6, 5, 4))


def _abs_tet_volume(P, quad):
    """|det(v1, v2, v3)| / 6 of the tetrahedron on the local vertices ``quad``;
    ``P[k]`` = coordinates (3, n_elem) of local vertex k.  The cross product is
    written out with NumPy's own operation order (product, product, difference:
    no fused multiply-add), so the bits equal ``np.cross``'s."""
    i0, i1, i2, i3 = quad
    v1 = P[i1] - P[i0]
    v2 = P[i2] - P[i0]
    v3 = P[i3] - P[i0]
    c0 = v1[1] * v2[2] - v1[2] * v2[1]
    c1 = v1[2] * v2[0] - v1[0] * v2[2]
    c2 = v1[0] * v2[1] - v1[1] * v2[0]
    return np.abs(c0 * v3[0] + c1 * v3[1] + c2 * v3[2]) / 6.0


def _hex_volumes_native(t_conn, p_coords):
    """The same sum in multi-threaded C++ (``sktb_host_hex_volumes``, csrc/host_setup.cu:
    host pointers, same operation order, no fused multiply-add: bit-identical); None
    when the library is not built."""
    import ctypes as C
    try:
        lib = _lib_mod.load()
    except Exception:
        return None
    t = np.ascontiguousarray(t_conn, dtype=np.int32)
    p = np.ascontiguousarray(p_coords, dtype=np.float64)
    if t.ndim != 2 or t.shape[0] != 8 or p.shape[0] != 3:
        return None
    vol = np.empty(t.shape[1])
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    if lib.sktb_host_hex_volumes(t.shape[1], p.shape[1], ptr(t), ptr(p), ptr(vol)) != 0:
        return None
    return vol


def _get_elements_volume_hex(t_conn, p_coords) -> np.ndarray:
    """Literal restatement of the reference's six-tetrahedra sum on the local
    index quadruples of ``fea/composer.py:191-248`` (under
    skfem's local vertex order the fifth term is degenerate; kept as is).  The
    coordinates of the seven local vertices the quadruples use are gathered once."""
    vol = _hex_volumes_native(t_conn, p_coords)
    if vol is not None:
        return vol
    used = sorted({k for quad in _HEX_TETS for k in quad})
    P = {k: p_coords[:, t_conn[k]] for k in used}
    vol = np.zeros(t_conn.shape[1])
    for quad in _HEX_TETS:
        vol += _abs_tet_volume(P, quad)
    return vol


def _get_elements_volume_tet(t_conn, p_coords) -> np.ndarray:
    """Signed det/6 (``fea/composer.py:164-180``); raises below -1e-12."""
    v1 = p_coords[:, t_conn[1]] - p_coords[:, t_conn[0]]
    v2 = p_coords[:, t_conn[2]] - p_coords[:, t_conn[0]]
    v3 = p_coords[:, t_conn[3]] - p_coords[:, t_conn[0]]
    c = np.cross(v1, v2, axis=0)
    vol = (c[0] * v3[0] + c[1] * v3[1] + c[2] * v3[2]) / 6.0
    bad = np.nonzero(vol < -1e-12)[0]
    if bad.size:
        print("Element", int(bad[0]), "has negative volume:", vol[bad[0]])
        raise ValueError("!!!")
    return vol


def get_elements_volume(mesh) -> np.ndarray:
    if isinstance(mesh, MeshTet):
        return _get_elements_volume_tet(mesh.t, mesh.p)
    if isinstance(mesh, MeshHex):
        return _get_elements_volume_hex(mesh.t, mesh.p)
    raise NotImplementedError("MeshTet or MeshHex")


# ------------------------------------------------------------------ assembly
def _csr_to_scipy(n, row_ptr, col_idx, vals):
    return sp.csr_matrix(
        (vals.cpu().numpy(), col_idx.cpu().numpy(), row_ptr.cpu().numpy()),
        shape=(n, n),
    )


def assemble_stiffness_matrix(basis, rho, E0: float, Emin: float, p: float,
                              nu: float, elem_func: Callable = simp_interpolation):
    """Global SIMP-weighted elasticity matrix as a SciPy CSR matrix
    (reference ``fea/composer.py:53-101``); computed on the GPU."""
    from sktopt._b200 import device as dev
    dm = dev.device_mesh(basis.mesh)
    ke0 = dm.unit_ke(0, basis.X, basis.W, nu=nu)
    E = dev.interpolate_modulus(dev.to_dev(rho), E0, Emin, p, ramp=is_ramp(elem_func))
    vals = dm.assemble(3, ke0, scale=E)
    rp, ci = dm.dof_pattern(3)
    return _csr_to_scipy(3 * dm.n_nodes, rp, ci, vals)


def assemble_conduction_matrix(basis, rho, k0: float, kmin: float, p: float,
                               elem_func: Callable = simp_interpolation):
    """Global SIMP-weighted conduction matrix (``fea/composer.py:104-145``)."""
    from sktopt._b200 import device as dev
    dm = dev.device_mesh(basis.mesh)
    ke0 = dm.unit_ke(1, basis.X, basis.W)
    k = dev.interpolate_modulus(dev.to_dev(rho), k0, kmin, p, ramp=is_ramp(elem_func))
    vals = dm.assemble(1, ke0, scale=k)
    rp, ci = dm.dof_pattern(1)
    return _csr_to_scipy(dm.n_nodes, rp, ci, vals)


def assemble_stiffness_matrix_simp(basis, rho, E0, Emin, p, nu):
    return assemble_stiffness_matrix(basis, rho, E0, Emin, p, nu, simp_interpolation)


def assemble_stiffness_matrix_ramp(basis, rho, E0, Emin, p, nu):
    return assemble_stiffness_matrix(basis, rho, E0, Emin, p, nu, ramp_interpolation)


# ----------------------------------------------- remaining public names -----
def adjacency_matrix(mesh) -> list:
    """Face neighbours of every tetrahedron as a list of lists (reference
    ``fea/composer.py:351-371``, a dict-of-faces Python loop there; vectorised
    here with the mesh's facet tables)."""
    f2t = np.asarray(mesh.f2t)                     # (2, n_facets), -1 = boundary
    both = f2t[:, (f2t >= 0).all(axis=0)]
    adjacency = [[] for _ in range(mesh.nelements)]
    for i, j in both.T.tolist():
        adjacency[i].append(j)
        adjacency[j].append(i)
    return adjacency


def strain_energy_skfem(basis, rho, u, E0, Emin, p, nu, elem_func=simp_interpolation):
    """Element strain energies of one displacement field (``fea/composer.py:402-423``;
    the same function also lives in ``fea/solver_elastic.py``, where it is built)."""
    from sktopt.fea import solver_elastic
    return solver_elastic.strain_energy_skfem(basis, rho, u, E0, Emin, p, nu, elem_func)


def strain_energy_skfem_multi(basis, rho, U, E0, Emin, p, nu, elem_func=simp_interpolation):
    from sktopt.fea import solver_elastic
    return solver_elastic.strain_energy_skfem_multi(basis, rho, U, E0, Emin, p, nu, elem_func)


# ------------------------------------------------------- stress post-processing --
def compute_element_stress_tensor(w):
    """sigma = 2 mu eps(u) + lam tr(eps(u)) I at the quadrature points (reference
    ``fea/composer.py:444-466``, there a scikit-fem ``Functional`` body).  ``w``
    maps 'uh' to an object whose ``.grad`` is the (3, 3, n_elem, n_qp) displacement
    gradient (or to that array itself) and 'mu_elem' / 'lam_elem' to (n_qp, n_elem)
    arrays, as the reference passes them.  Returns (3, 3, n_elem, n_qp)."""
    uh = w["uh"]
    grad = np.asarray(getattr(uh, "grad", uh), dtype=np.float64)
    sym = 0.5 * (grad + np.swapaxes(grad, 0, 1))
    tr = np.trace(sym, axis1=0, axis2=1)
    mu = np.asarray(w["mu_elem"]).T[None, None, :, :]
    lam = np.asarray(w["lam_elem"]).T[None, None, :, :]
    return 2.0 * mu * sym + lam * np.eye(3)[:, :, None, None] * tr[None, None, :, :]


def stress_tensor_skfem(basis, rho, u, E0: float, Emin: float, p: float, nu: float,
                        elem_func: Callable = simp_interpolation):
    """Stress tensor of the displacement field ``u`` at every quadrature point of
    ``basis``: array (3, 3, n_elem, n_qp), the layout
    ``von_mises_from_stress_tensor`` consumes (reference ``:469-494``).

    (The reference pushes the tensor through ``Functional.elemental``, whose sum
    over axis 1 collapses the second tensor index instead of the quadrature axis;
    that accident is not reproduced -- the point values are returned.)"""
    from sktopt._b200 import device as dev
    from sktopt._b200 import lib as _lib
    import torch
    dev.require_cuda()
    dm = dev.device_mesh(basis.mesh)
    on_dev = isinstance(u, torch.Tensor) and u.is_cuda
    u_d = u if on_dev else dev.to_dev(np.ascontiguousarray(u, dtype=np.float64))
    rho_d = rho if (isinstance(rho, torch.Tensor) and rho.is_cuda) else dev.to_dev(rho)
    E = dev.interpolate_modulus(rho_d, E0, Emin, p, ramp=is_ramp(elem_func))
    N, G, dx, _ = dm.geom_tables(basis.X, basis.W)
    nq = int(basis.X.shape[1])
    out = torch.empty((3, 3, dm.n_elem, nq), dtype=dev.F64, device="cuda")
    _lib.check(_lib.load().sktb_element_stress(
        dm.handle, nq, dev._ptr(dm.elem_class), dev._ptr(G), dev._ptr(E), float(nu),
        dev._ptr(u_d), dev._ptr(out), dev._stream()))
    return out if on_dev else out.cpu().numpy()


def von_mises_from_stress_tensor(stress_tensor):
    """Von Mises stress (n_elem, n_qp) from a (3, 3, n_elem, n_qp) stress tensor
    (reference ``:497-519``); NumPy arrays and CUDA tensors are both accepted."""
    s = stress_tensor
    sqrt = np.sqrt
    try:
        import torch
        if isinstance(s, torch.Tensor):
            sqrt = torch.sqrt
    except ImportError:            # pragma: no cover
        pass
    s_xx, s_yy, s_zz = s[0, 0], s[1, 1], s[2, 2]
    s_xy, s_yz, s_zx = s[0, 1], s[1, 2], s[2, 0]
    return sqrt(0.5 * ((s_xx - s_yy) ** 2 + (s_yy - s_zz) ** 2 + (s_zz - s_xx) ** 2
                       + 6.0 * (s_xy ** 2 + s_yz ** 2 + s_zx ** 2)))
