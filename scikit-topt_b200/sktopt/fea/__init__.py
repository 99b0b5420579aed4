from sktopt.fea.solver_elastic import FEM_SimpLinearElasticity
from sktopt.fea.solver_heat import FEM_SimpLinearHeatConduction

__all__ = ["FEM_SimpLinearElasticity", "FEM_SimpLinearHeatConduction"]
