"""Options shim for the reference's optional PETSc backend
(``fea/_petsc_compat.py``, ``fea/solver_petsc.py:12``).  PETSc is an alternative
CPU backend and out of scope; the dataclass is kept so that configurations
written for the reference still construct.  Selecting ``"petsc"`` raises."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class PETScOptions:
    ksp_type: str = "cg"
    pc_type: str = "gamg"
    pc_factor_mat_solver_type: str | None = None
    options_prefix: str | None = None
    use_set_from_options: bool = True


def _missing_petsc(*_args, **_kwargs):
    raise RuntimeError(
        "PETSc support is unavailable in the B200 build; use the default "
        "device PCG solver.")


def normalize_petsc_options(petsc_options, direct: bool = False) -> PETScOptions:
    if isinstance(petsc_options, PETScOptions):
        return petsc_options
    base = PETScOptions(ksp_type="preonly" if direct else "cg",
                        pc_type="lu" if direct else "gamg")
    if not isinstance(petsc_options, dict):
        return base
    opt = lambda k: None if petsc_options.get(k) is None else str(petsc_options[k])
    return PETScOptions(
        ksp_type=str(petsc_options.get("ksp_type", base.ksp_type)),
        pc_type=str(petsc_options.get("pc_type", base.pc_type)),
        pc_factor_mat_solver_type=opt("pc_factor_mat_solver_type"),
        options_prefix=opt("options_prefix"),
        use_set_from_options=bool(
            petsc_options.get("use_set_from_options", base.use_set_from_options)),
    )


def petsc_options_for_solver(chosen_solver: str, petsc_options):
    if chosen_solver in {"petsc", "petsc_spdirect"}:
        return _missing_petsc()
    return None if petsc_options is None else normalize_petsc_options(petsc_options)


solve_u_petsc = _missing_petsc
solve_u_petsc_multi = _missing_petsc
