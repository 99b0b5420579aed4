"""B200-native build of scikit-topt's per-iteration FEA + sensitivity loop.

Drop-in for the hot path of ``sktopt`` 0.3.9: ``sktopt.mesh`` task definitions,
``sktopt.fea`` solver entry points, ``sktopt.filters`` and the OC / LogMOC
optimisers in ``sktopt.core`` keep the reference's names and signatures and run
on hand-written sm_100a kernels (``csrc/``) through the C ABI in
``include/sktopt_b200.h``.
"""
from . import tools, fea, filters, mesh, core

__version__ = "0.3.9+b200.1"

__all__ = ["__version__", "mesh", "core", "fea", "tools", "filters"]
