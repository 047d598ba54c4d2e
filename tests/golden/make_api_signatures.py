"""Records the public call signatures of the reference's hot-path modules (parameter names, order, defaults) as a small JSON
fixture, read with `ast` from /root/reference (nothing is imported or executed).  tests/test_host.py checks the drop-in's
classes and functions against it: same names, same order, same defaults; the drop-in may only ADD trailing keyword parameters.
usage: python tests/golden/make_api_signatures.py [/root/reference]"""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FILES = {"ConvectionDiffusion_Solver": "Solvers/ConvectionDiffusion_Solver.py", "NavierStokes_Solver": "Solvers/NavierStokes_Solver.py",
         "SEM": "Solvers/SEM.py", "GLL": "Solvers/GLL.py", "Boussinesq_SequentialCoupler": "OpenMDAO/Boussinesq_SequentialCoupler.py"}


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    out = [{"name": n, "default": d} for n, d in zip(names, defaults)]
    out += [{"name": x.arg, "default": ast.unparse(d) if d is not None else None, "kwonly": True}
            for x, d in zip(a.kwonlyargs, a.kw_defaults)]
    return out


api = {}
for mod, rel in FILES.items():
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    entry = {"functions": {}, "classes": {}}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            entry["functions"][node.name] = sig(node)
        elif isinstance(node, ast.ClassDef):
            entry["classes"][node.name] = {m.name: sig(m) for m in node.body if isinstance(m, ast.FunctionDef)}
    api[mod] = entry
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")
json.dump(api, open(out, "w"), indent=1, sort_keys=True)
print(out, {m: (len(e["functions"]), {c: len(v) for c, v in e["classes"].items()}) for m, e in api.items()})
