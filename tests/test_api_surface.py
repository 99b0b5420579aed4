"""Drop-in boundary: every public function, class and method the
reference defines in its hot-path modules exists in this package under the same
name with the same leading positional parameters.  The reference surface was
extracted by parsing its sources (tests/golden/make_api_surface.py ->
tests/golden/api_surface.json); names that are deliberately not built are listed
here with the reason."""
import importlib
import inspect
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

OUT_OF_SCOPE = {
    "core.misc.add_common_arguments": "argparse wiring of the reference's CLI examples (DESIGN.md 8)",
    "core.misc.args2OC_Config_dict": "argparse wiring of the reference's CLI examples",
    "fea.solver_heat.avg_temp_skfem": "elemental (T - T_env) integrals: unused by the optimiser, "
                                      "which takes J = sum(T) (solver_heat.py:871-873)",
    "fea.solver_heat.avg_temp_skfem_multi": "same",
    "fea.solver_heat.get_robin_virtual": "returns skfem form objects; the virtual Robin terms are "
                                         "kernels here (sktb_robin_virtual_scale / assemble_terms)",
    "fea.solver_heat.heat_exchange_objective": "takes a skfem CellBasis; the objective is evaluated "
                                               "by FEM_SimpLinearHeatConduction (sktb_heat_exchange_local)",
    "fea.solver_heat.solve_heat_system_multi": "LU / PETSc driver on a SciPy matrix; the enforced "
                                               "solve is FEM_SimpLinearHeatConduction._solve_enforced",
    "fea.solver_heat.solve_multi_load": "same (assembly + LU on SciPy matrices)",
    "fea.solver_heat.solve_scipy": "same",
    "fea.solver_heat.solve_scipy_heat_multi_enforce": "same",
    "filters.helmholtz_filter_nodal.test_main": "a demo entry point, not API",
    "mesh.toy_problem.load_mesh_auto": "meshio loader (.msh), out of scope",
    "tools.history.compare_histories_data_and_plot_type": "matplotlib comparison plot",
}

# accepted signature differences
SIGNATURE_NOTES = {
    "core.optimizers.common_density.DensityMethod.rho_update":
        "the reference's base-class declaration is stale; its call site (common_density.py:1112-1129) "
        "and both optimisers (oc.py:149-167, logmoc.py:90-108) use the 16-argument form kept here",
    "core.optimizers.common_density.DensityMethodBase.rho_update": "same",
    "tools.timer.SectionTimer.plot": "matplotlib plots raise: arguments are accepted as *args",
    "tools.timer.SectionTimer.plot_bar": "same",
    "tools.timer.SectionTimer.plot_pie": "same",
    "tools.timer.SectionTimer.save_plot": "no-op without matplotlib: arguments accepted as *args",
}


def _positional(f):
    try:
        sig = inspect.signature(f)
    except (TypeError, ValueError):
        return None
    return [p.name for p in sig.parameters.values()
            if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]


def test_reference_api_surface_is_present():
    with open(os.path.join(HERE, "golden", "api_surface.json")) as f:
        surface = json.load(f)
    missing, differs, checked = [], [], 0
    for mod, m in surface.items():
        M = importlib.import_module("sktopt." + mod)
        for fn, sg in m["functions"].items():
            name = f"{mod}.{fn}"
            if not hasattr(M, fn):
                missing.append(name)
                continue
            checked += 1
            mine = _positional(getattr(M, fn))
            if mine is not None and mine[:len(sg["args"])] != sg["args"]:
                differs.append(name)
        for cn, c in m["classes"].items():
            if not hasattr(M, cn):
                missing.append(f"{mod}.{cn}")
                continue
            C = getattr(M, cn)
            for meth, sg in c["methods"].items():
                name = f"{mod}.{cn}.{meth}"
                if not hasattr(C, meth):
                    missing.append(name)
                    continue
                checked += 1
                static = inspect.getattr_static(C, meth)
                if isinstance(static, property):
                    continue
                ref = sg["args"][1:] if isinstance(static, classmethod) else sg["args"]
                mine = _positional(getattr(C, meth))
                if mine is not None and mine[:len(ref)] != ref:
                    differs.append(name)
    assert checked >= 200
    assert sorted(missing) == sorted(OUT_OF_SCOPE), sorted(set(missing) ^ set(OUT_OF_SCOPE))
    assert sorted(differs) == sorted(SIGNATURE_NOTES), sorted(set(differs) ^ set(SIGNATURE_NOTES))


def test_dataclass_fields_of_tasks_and_configs():
    """FEMDomain's 22 fields and the optimiser config fields keep the reference names."""
    import dataclasses
    with open(os.path.join(HERE, "golden", "api_surface.json")) as f:
        surface = json.load(f)
    for mod, cls in (("mesh.task_common", "FEMDomain"), ("mesh.task_elastic", "LinearElasticity"),
                     ("mesh.task_heat", "LinearHeatConduction"),
                     ("core.optimizers.common_density", "DensityMethodConfig"),
                     ("core.optimizers.oc", "OC_Config"), ("core.optimizers.logmoc", "LogMOC_Config")):
        ref_fields = surface[mod]["classes"][cls]["fields"]
        C = getattr(importlib.import_module("sktopt." + mod), cls)
        mine = {f.name for f in dataclasses.fields(C)}
        lacking = [f for f in ref_fields if f not in mine]
        assert not lacking, (cls, lacking)
