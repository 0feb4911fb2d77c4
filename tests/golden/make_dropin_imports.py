"""List every name the reference's entry scripts take from the `crowdsam` / `segment_anything_cs` packages
(tools/test.py, tools/batch_eval.py, tools/demo.py; plus what crowdsam/model.py itself imports from
segment_anything_cs, i.e. the surface a replacement predictor package has to offer).
Run in the build container:  python tests/golden/make_dropin_imports.py  ->  tests/golden/dropin_imports.json"""
import ast
import json
import os
import sys

REF = os.environ.get("CROWDSAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
PKGS = ("crowdsam", "segment_anything_cs")


def scan(path):
    tree = ast.parse(open(path).read())
    out, aliases = set(), {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in PKGS:
            for a in node.names:
                out.add((node.module, a.name))
        elif isinstance(node, ast.Import):
            for a in node.names:
                if a.name.split(".")[0] in PKGS:
                    out.add((a.name, ""))
                    aliases[a.asname or a.name] = a.name
    for node in ast.walk(tree):      # attribute uses through a module alias: utils.load_config(...)
        if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and node.value.id in aliases:
            out.add((aliases[node.value.id], node.attr))
    return sorted(out)


def collect():
    res = {}
    for rel in ("tools/test.py", "tools/batch_eval.py", "tools/demo.py"):
        res[rel] = scan(os.path.join(REF, rel))
    # the predictor-package surface crowdsam/model.py consumes (only segment_anything_cs names)
    res["crowdsam/model.py"] = [x for x in scan(os.path.join(REF, "crowdsam/model.py")) if x[0].startswith("segment_anything_cs")]
    return res


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs the reference tree")
    with open(os.path.join(HERE, "dropin_imports.json"), "w") as f:
        json.dump(collect(), f, indent=1, sort_keys=True)
    print(json.dumps(collect(), indent=1))
