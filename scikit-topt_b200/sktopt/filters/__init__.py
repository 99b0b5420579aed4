"""Density filters of the path (``sktopt.filters``): the nodal Helmholtz (PDE)
filter and the neighbour-weighted ("spacial", the reference's spelling) filter.
``HelmholtzFilterElement`` is dead code in the reference (0.3.9) and not built."""
from .helmholtz_filter_nodal import HelmholtzFilterNodal
from .spacial import SpacialFilter

__all__ = ["HelmholtzFilterNodal", "SpacialFilter"]

for _cls in (HelmholtzFilterNodal, SpacialFilter):
    _cls.__module__ = __name__
del _cls
