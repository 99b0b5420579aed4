"""Material-interpolation derivatives and compliance sensitivities
(reference ``core/derivatives.py:25-89``).

dC/drho = -2 U_e (dE/drho) / max(E, 1e-12), with rho clamped to >= 1e-6 inside
the SIMP power (:26,:39) -- while the assembly uses the unclamped rho^p
(SURVEY.md B-5).  CUDA tensors are routed to the ``sktb_dc_drho`` kernel."""
import os

import numpy as np


def _is_dev(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _sensitivity_mode() -> str:
    raw = os.environ.get("SCITOPT_SENSITIVITY_MODE", "current")
    if raw.strip().lower() not in ("current", "default", "physical", "scaled"):
        raise ValueError(
            "SCITOPT_SENSITIVITY_MODE must be one of 'current', 'default', "
            f"'physical', or 'scaled', got: {raw}")
    return "current"


def _E_simp(rho, E0, Emin, p):
    return Emin + (E0 - Emin) * np.maximum(rho, 1e-6) ** p


def _E_ramp(rho, E0, Emin, p):
    return Emin + (E0 - Emin) * (rho / (1.0 + p * (1.0 - rho)))


def dE_drho_simp(rho, E0, Emin, p):
    return p * (E0 - Emin) * np.maximum(rho, 1e-6) ** (p - 1)


def dE_drho_ramp(rho, E0, Emin, p):
    denom = 1.0 + p * (1.0 - rho)
    return (E0 - Emin) * (denom - p * rho) / (denom ** 2)


def _dC(rho, strain_energy, E0, Emin, p, ramp):
    if _is_dev(rho):
        from sktopt._b200 import device as dev
        return dev.dc_drho(rho, strain_energy.contiguous(), E0, Emin, p, ramp=ramp)
    dE = (dE_drho_ramp if ramp else dE_drho_simp)(rho, E0, Emin, p)
    E = (_E_ramp if ramp else _E_simp)(rho, E0, Emin, p)
    return -2.0 * strain_energy * dE / np.maximum(E, 1e-12)


def dC_drho_simp(rho, strain_energy, E0, Emin, p):
    return _dC(rho, strain_energy, E0, Emin, p, False)


def dC_drho_ramp(rho, strain_energy, E0, Emin, p):
    return _dC(rho, strain_energy, E0, Emin, p, True)


def dE_drho_ramp_inplace(rho, out, E0, Emin, p):
    np.copyto(out, dE_drho_ramp(rho, E0, Emin, p))


def dC_drho_ramp_inplace(rho, strain_energy, out, E0, Emin, p):
    np.copyto(out, dC_drho_ramp(rho, strain_energy, E0, Emin, p))
