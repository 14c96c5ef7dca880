"""Extract the public call signatures of the reference's wflib/IDEAL_model.py with `ast` (no TensorFlow import)
into tests/golden/reference_signatures.json.  Run in the build container only.  TEST INFRASTRUCTURE ONLY."""
import ast
import json
import os

REF = os.environ.get("IDEALGAN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [ast.unparse(d) for d in a.defaults]
    pad = [None] * (len(names) - len(defaults)) + defaults
    return [[n, d] for n, d in zip(names, pad)]


def main():
    tree = ast.parse(open(os.path.join(REF, "wflib", "IDEAL_model.py")).read())
    out = {"functions": {}, "classes": {}, "constants": []}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            out["functions"][node.name] = sig(node)
        elif isinstance(node, ast.ClassDef):
            out["classes"][node.name] = {m.name: sig(m) for m in node.body if isinstance(m, ast.FunctionDef)}
        elif isinstance(node, ast.Assign):
            out["constants"] += [t.id for t in node.targets if isinstance(t, ast.Name)]
    out["constants"] = sorted(set(out["constants"]))
    path = os.path.join(ROOT, "tests", "golden", "reference_signatures.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, len(out["functions"]), "functions", len(out["classes"]), "classes")


if __name__ == "__main__":
    main()
