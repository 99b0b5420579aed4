from sktopt.core.optimizers.common_density import DensityMethodConfig
from sktopt.core.optimizers.common_density import DensityMethod_OC_Config
from sktopt.core.optimizers.common_density import DensityMethod
from sktopt.core.optimizers.common_density import DensityState
from sktopt.core.optimizers.oc import OC_Config, OC_Optimizer
from sktopt.core.optimizers.logmoc import LogMOC_Config, LogMOC_Optimizer
from sktopt.core import optimizers, projection, derivatives, misc

# the reference aliases its "scaled" variants to the plain ones when absent
OCScaled_Config, OCScaled_Optimizer = OC_Config, OC_Optimizer
LogMOCScaled_Config, LogMOCScaled_Optimizer = LogMOC_Config, LogMOC_Optimizer

for _c in (DensityMethodConfig, DensityMethod_OC_Config, DensityMethod,
           DensityState, OC_Config, OC_Optimizer, LogMOC_Config, LogMOC_Optimizer):
    _c.__module__ = __name__

__all__ = [
    "DensityMethodConfig", "DensityMethod_OC_Config", "DensityMethod",
    "DensityState", "OC_Config", "OC_Optimizer", "OCScaled_Config",
    "OCScaled_Optimizer", "LogMOC_Config", "LogMOC_Optimizer",
    "LogMOCScaled_Config", "LogMOCScaled_Optimizer",
]
