"""Host-side mesh / basis substrate replacing the scikit-fem surface the
reference uses; scikit-fem is not installed here."""
from .mesh import Mesh, MeshHex, MeshTet
from .basis import (
    Basis, ElementHex1, ElementTetP1, ElementVector, DofsView,
    hex_quadrature, tet_quadrature,
)
from .facet import facet_area, facet_load, facet_mass, facet_quadrature

__all__ = [
    "Mesh", "MeshHex", "MeshTet", "Basis", "ElementHex1", "ElementTetP1",
    "ElementVector", "DofsView", "hex_quadrature", "tet_quadrature",
    "facet_area", "facet_load", "facet_mass", "facet_quadrature",
]
