"""Density filters of the path (``sktopt.filters``): the nodal Helmholtz (PDE)
filter and the neighbour-weighted ("spacial", the reference's spelling) filter.
``HelmholtzFilterElement`` (element-graph Helmholtz filter) is importable from
``sktopt.filters.helmholtz_filter_element``; like the reference (0.3.9) the
package namespace does not export it."""
from .helmholtz_filter_nodal import HelmholtzFilterNodal
from .spacial import SpacialFilter

__all__ = ["HelmholtzFilterNodal", "SpacialFilter"]

for _cls in (HelmholtzFilterNodal, SpacialFilter):
    _cls.__module__ = __name__
del _cls
