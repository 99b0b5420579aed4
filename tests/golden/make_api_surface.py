#!/usr/bin/env python
"""Extracts the public API surface of the reference's hot-path modules (names
and call signatures only, by parsing the source with `ast`: nothing is imported
or copied) into tests/golden/api_surface.json.  Run in the build container, where
/root/reference exists:  python tests/golden/make_api_surface.py"""
import ast
import json
import os

REF = "/root/reference/scikit-topt/sktopt"
MODULES = [
    "mesh/task_common.py", "mesh/task_elastic.py", "mesh/task_heat.py", "mesh/toy_problem.py",
    "mesh/utils.py", "fea/composer.py", "fea/solver.py", "fea/solver_elastic.py",
    "fea/solver_heat.py", "filters/base.py", "filters/helmholtz_filter_nodal.py",
    "filters/spacial.py", "core/derivatives.py", "core/projection.py", "core/misc.py",
    "core/optimizers/common_density.py", "core/optimizers/oc.py", "core/optimizers/logmoc.py",
    "tools/scheduler.py", "tools/history.py", "tools/timer.py",
]


def sig(fn: ast.FunctionDef):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    return {"args": pos, "n_defaults": len(a.defaults), "kwonly": [x.arg for x in a.kwonlyargs],
            "vararg": bool(a.vararg), "kwarg": bool(a.kwarg)}


def main():
    out = {}
    for rel in MODULES:
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        mod = {"functions": {}, "classes": {}}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
                mod["functions"][node.name] = sig(node)
            elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
                methods, fields = {}, []
                for it in node.body:
                    if isinstance(it, ast.FunctionDef) and (not it.name.startswith("_")
                                                            or it.name == "__init__"):
                        methods[it.name] = sig(it)
                    elif isinstance(it, ast.AnnAssign) and isinstance(it.target, ast.Name):
                        fields.append(it.target.id)
                mod["classes"][node.name] = {"methods": methods, "fields": fields}
        out[rel[:-3].replace("/", ".")] = mod
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(path, sum(len(m["functions"]) + len(m["classes"]) for m in out.values()), "top-level names")


if __name__ == "__main__":
    main()
