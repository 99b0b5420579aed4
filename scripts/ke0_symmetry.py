"""Groundwork for the next grid-operator kernel: the
element matrix Ke0 of a brick (trilinear hexahedron, isotropic elasticity) commutes with
the three reflections of the brick, so in the basis of their common eigenvectors -- a
Hadamard transform of the 8 corners, with the sign of a displacement component flipped
by the reflection along its own axis -- it is block diagonal: 8 blocks of 3 x 3.
y_e = Ke0 u_e then costs 2 x 72 additions + 72 FMA instead of 576 FMA.

CPU only (NumPy + the oracle's element matrix); prints the block structure and the
operation counts.  Usage: python scripts/ke0_symmetry.py [hx hy hz]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fem  # noqa: E402


def block_structure(h):
    """(T, B): the orthogonal symmetry-adapted basis T (24 x 24) and B = T^T Ke0 T for a
    brick of edge lengths h; B is block diagonal (8 blocks of 3 x 3) to rounding."""
    sys.path.insert(0, os.path.join(ROOT, "scikit-topt_b200"))
    from sktopt._fem import MeshHex
    cell = MeshHex.init_tensor([0.0, h[0]], [0.0, h[1]], [0.0, h[2]])
    p, t = cell.p, cell.t
    lam, mu = 0.3 / ((1 + 0.3) * (1 - 0.6)), 1.0 / (2 * (1 + 0.3))
    Ke = fem.elasticity_ke(p, t, np.array([lam]), np.array([mu]), intorder=2)[0]
    Ke = 0.5 * (Ke + Ke.T)
    # which corner is each local vertex?  (bits of the corner along x, y, z)
    bits = (p.T[t[:, 0]] > 0.5 * np.array(h)).astype(int)                       # (8, 3)
    # characters of Z2^3: chi_s(corner) = (-1)^(s . bits); component c picks up an extra
    # factor from the reflection along its own axis, i.e. it lives in the character s ^ e_c
    T = np.zeros((24, 24))
    for s in range(8):
        sb = np.array([(s >> 0) & 1, (s >> 1) & 1, (s >> 2) & 1])
        for c in range(3):
            sc = sb.copy()
            sc[c] ^= 1
            chi = (-1.0) ** (bits @ sc)
            T[3 * np.arange(8) + c, 3 * s + c] = chi / np.sqrt(8.0)
    return T, T.T @ Ke @ T, Ke


def off_block_ratio(h):
    T, B, Ke = block_structure(h)
    off = B.copy()
    for s in range(8):
        off[3 * s:3 * s + 3, 3 * s:3 * s + 3] = 0.0
    return float(np.abs(off).max() / np.abs(Ke).max()), float(np.abs(T.T @ T - np.eye(24)).max())


if __name__ == "__main__":
    h = [float(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else [0.0577, 0.0577, 0.0571]
    T, B, Ke = block_structure(h)
    print("brick %s: max |off-block| / max |Ke| = %.2e" % (h, off_block_ratio(h)[0]))
    for s in range(8):
        blk = B[3 * s:3 * s + 3, 3 * s:3 * s + 3]
        print("block", s, "eigenvalues", np.round(np.linalg.eigvalsh(blk), 6))
    print("rigid-body modes -> the six zero eigenvalues; dense product: 576 FMA per element;")
    print("transformed: 3 components x 2 transforms x 24 additions + 8 blocks x 9 FMA = 216")
