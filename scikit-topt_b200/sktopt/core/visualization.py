"""VTU export of a mesh with point / cell fields, without meshio.

The reference writes its per-tick ``info_mesh-XXXXXXXX.vtu`` (cell fields
``rho_projected`` and ``energy``) and the task's ``condition.vtu`` through
``meshio.Mesh(...).write`` (``core/visualization.py:22-84``,
``mesh/task_common.py:360-396``).  meshio is not a dependency here: the VTK XML
UnstructuredGrid format is written directly (inline base64 "binary" DataArrays
with a UInt64 byte-count header, uncompressed, little endian), which ParaView,
PyVista and meshio read.

Like the reference, the connectivity is ``mesh.t.T`` as it stands, i.e. in
scikit-fem's local vertex order (the reference hands ``mesh.t.T`` to meshio
without ``skfem``'s own to-VTK permutation), so files are interchangeable with
the reference's, quirk included.

Rendering helpers of the reference (PyVista screenshots, matplotlib histograms,
GIF assembly) are I/O side-cars outside the hot path; they are not rebuilt.
"""
from __future__ import annotations

import base64
import os
import re
from typing import Optional

import numpy as np

_VTK_TYPE = {8: 12, 4: 10}          # hexahedron, tetra
_VTK_NAME = {np.dtype("float64"): "Float64", np.dtype("float32"): "Float32",
             np.dtype("int64"): "Int64", np.dtype("int32"): "Int32",
             np.dtype("uint8"): "UInt8"}


class VisualizationUnavailableWarning(UserWarning):
    pass


def _b64(a: np.ndarray) -> str:
    raw = np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
    head = np.array([len(raw)], dtype="<u8").tobytes()
    return base64.b64encode(head).decode("ascii") + base64.b64encode(raw).decode("ascii")


def _data_array(name: str, a: np.ndarray, ncomp: int = 1) -> str:
    a = np.asarray(a)
    if a.dtype == np.bool_:
        a = a.astype(np.uint8)
    if a.dtype not in _VTK_NAME:
        a = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
    comp = f' NumberOfComponents="{ncomp}"' if ncomp > 1 else ""
    return (f'<DataArray type="{_VTK_NAME[a.dtype]}" Name="{name}"{comp} format="binary">'
            f"{_b64(a.ravel())}</DataArray>\n")


def export_mesh_with_info(
    mesh,
    point_data_values: Optional[list] = None,
    point_data_names: Optional[list] = None,
    cell_data_values: Optional[list] = None,
    cell_data_names: Optional[list] = None,
    filepath: str = "output.vtu",
):
    """Write ``mesh`` (``MeshHex`` / ``MeshTet``: ``p`` (3, n_nodes), ``t``
    (nen, n_elem)) and its fields to a ``.vtu`` file.  Same arguments as the
    reference's function of this name."""
    nen, ne = mesh.t.shape
    if nen not in _VTK_TYPE:
        raise ValueError(f"Unsupported mesh type: {type(mesh)}")
    nn = mesh.p.shape[1]
    pts = np.zeros((nn, 3), dtype=np.float64)
    pts[:, :mesh.p.shape[0]] = mesh.p.T
    out = ['<?xml version="1.0"?>\n'
           '<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian" '
           'header_type="UInt64">\n<UnstructuredGrid>\n'
           f'<Piece NumberOfPoints="{nn}" NumberOfCells="{ne}">\n<Points>\n',
           _data_array("Points", pts, 3), "</Points>\n<Cells>\n",
           _data_array("connectivity", np.ascontiguousarray(mesh.t.T, dtype=np.int64)),
           _data_array("offsets", nen * np.arange(1, ne + 1, dtype=np.int64)),
           _data_array("types", np.full(ne, _VTK_TYPE[nen], dtype=np.uint8)),
           "</Cells>\n"]
    for tag, names, vals, n in (("PointData", point_data_names, point_data_values, nn),
                                ("CellData", cell_data_names, cell_data_values, ne)):
        if not (names and vals):
            continue
        out.append(f"<{tag}>\n")
        for name, v in zip(names, vals):
            v = np.asarray(v)
            if v.shape[0] != n:
                raise ValueError(f"{tag} '{name}' has {v.shape[0]} rows, expected {n}")
            out.append(_data_array(name, v, 1 if v.ndim == 1 else int(np.prod(v.shape[1:]))))
        out.append(f"</{tag}>\n")
    out.append("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    tmp = f"{filepath}.tmp{os.getpid()}"
    with open(tmp, "w") as f:
        f.write("".join(out))
    os.replace(tmp, filepath)


_NP_OF = {v: k for k, v in _VTK_NAME.items()}


def read_vtu(filepath: str) -> dict:
    """Minimal reader of the files ``export_mesh_with_info`` writes (used by the
    round-trip tests and for restarts from a field): returns ``points``,
    ``connectivity``, ``offsets``, ``types``, ``point_data`` and ``cell_data``."""
    txt = open(filepath).read()

    def arrays(section):
        m = re.search(rf"<{section}>(.*?)</{section}>", txt, re.S)
        res = {}
        if not m:
            return res
        for am in re.finditer(r'<DataArray type="(\w+)" Name="([^"]+)"(?: NumberOfComponents="(\d+)")?'
                              r' format="binary">([^<]*)</DataArray>', m.group(1)):
            ty, name, nc, payload = am.groups()
            raw = base64.b64decode(payload[12:])       # 8-byte header = 12 base64 chars
            a = np.frombuffer(raw, dtype=_NP_OF[ty].newbyteorder("<")).astype(_NP_OF[ty])
            res[name] = a.reshape(-1, int(nc)) if nc else a
        return res

    cells = arrays("Cells")
    return dict(points=arrays("Points")["Points"], connectivity=cells["connectivity"],
                offsets=cells["offsets"], types=cells["types"],
                point_data=arrays("PointData"), cell_data=arrays("CellData"))


def write_mesh_with_info_as_image(mesh_path: str, mesh_scalar_name: str, clim: tuple,
                                  image_path: str, image_title: str) -> bool:
    """PyVista off-screen rendering in the reference; not available here."""
    import warnings
    warnings.warn("image export needs pyvista; skipped", VisualizationUnavailableWarning)
    return False
