from sktopt.filters.spacial import SpacialFilter
from sktopt.filters.helmholtz_filter_nodal import HelmholtzFilterNodal

SpacialFilter.__module__ = "sktopt.filters"
HelmholtzFilterNodal.__module__ = "sktopt.filters"

__all__ = ["SpacialFilter", "HelmholtzFilterNodal"]
