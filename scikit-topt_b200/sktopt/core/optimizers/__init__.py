from sktopt.core.optimizers.oc import OC_Config
from sktopt.core.optimizers.oc import OC_Optimizer
from sktopt.core.optimizers.logmoc import LogMOC_Config
from sktopt.core.optimizers.logmoc import LogMOC_Optimizer

__all__ = ["OC_Config", "OC_Optimizer", "LogMOC_Config", "LogMOC_Optimizer"]
