"""Task definitions of the path (``sktopt.mesh``): the boundary-condition
container ``FEMDomain`` and its two physics, plus the toy problems and the mesh
helpers.  The classes report ``sktopt.mesh`` as their module, like the reference's
package, so that pickles and reprs written against it keep resolving."""
from . import toy_problem, utils
from .task_common import FEMDomain
from .task_elastic import LinearElasticity
from .task_heat import LinearHeatConduction

__all__ = ["FEMDomain", "LinearElasticity", "LinearHeatConduction", "toy_problem", "utils"]

for _cls in (FEMDomain, LinearElasticity, LinearHeatConduction):
    _cls.__module__ = __name__
del _cls
