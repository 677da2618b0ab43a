"""The product package must never import, call or link the test oracle (or any CPU fallback)."""
import ast
import os

from conftest import ROOT


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffsensei_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                for n in names:
                    assert not (n == "oracle" or n.startswith("oracle.")), f"{f} imports {n}"
    for dirpath, _dirs, files in os.walk(os.path.join(pkg, "csrc")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), f
