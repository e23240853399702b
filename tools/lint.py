#!/usr/bin/env python
"""Self-contained lint gate (no flake8 / pyflakes in the image): the checks of the reference's pre-commit setup that
need no third-party tool (.pre-commit-config.yaml:1-47, .tools/codestyle/docstring_checker.py) --

  * every module has a docstring,
  * no line longer than 130 columns (185 under tests/ and tools/), no tabs, no trailing whitespace,
  * no unused imports (names imported but never referenced; ``# noqa`` lines and ``__init__`` re-exports are skipped),
  * no duplicate top-level definitions, no bare ``except:``, no mutable default arguments,
  * every file compiles.

    python tools/lint.py [paths...]        # exit code 1 if anything is found
"""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT = ["edl_b200", "examples", "tests", "tools", "k8s", "baseline", "bench.py", "__graft_entry__.py", "setup.py"]
MAX_COL = 130


def py_files(paths):
    for p in paths:
        p = os.path.join(ROOT, p) if not os.path.isabs(p) else p
        if os.path.isfile(p) and p.endswith(".py"):
            yield p
        for d, dirs, files in os.walk(p):
            dirs[:] = [x for x in dirs if x not in ("__pycache__", "_ref", "build")]
            for f in files:
                if f.endswith(".py"):
                    yield os.path.join(d, f)


def check(path):
    out = []
    src = open(path, encoding="utf-8").read()
    lines = src.split("\n")
    rel = os.path.relpath(path, ROOT)
    limit = 185 if rel.startswith(("tests", "tools")) else MAX_COL      # scripts and tests: long literals are fine
    for i, ln in enumerate(lines, 1):
        if len(ln) > limit:
            out.append((i, "line too long (%d > %d)" % (len(ln), limit)))
        if "\t" in ln:
            out.append((i, "tab character"))
        if ln != ln.rstrip():
            out.append((i, "trailing whitespace"))
    try:
        tree = ast.parse(src, path)
    except SyntaxError as e:
        return [(e.lineno or 0, "syntax error: %s" % e.msg)]
    if src.strip() and ast.get_docstring(tree) is None and os.path.basename(path) != "__init__.py":
        out.append((1, "module docstring missing"))
    noqa = {i for i, ln in enumerate(lines, 1) if "# noqa" in ln}
    imported = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                imported[(a.asname or a.name).split(".")[0]] = node.lineno
        elif isinstance(node, ast.ImportFrom):
            for a in node.names:
                if a.name != "*":
                    imported[a.asname or a.name] = node.lineno
        elif isinstance(node, ast.ExceptHandler) and node.type is None:
            out.append((node.lineno, "bare except"))
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef)):
            for d in node.args.defaults + [d for d in node.args.kw_defaults if d is not None]:
                if isinstance(d, (ast.List, ast.Dict, ast.Set)):
                    out.append((node.lineno, "mutable default argument in %s()" % node.name))
    used = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            used.add(node.id)
        elif isinstance(node, ast.Attribute):
            n = node
            while isinstance(n, ast.Attribute):
                n = n.value
            if isinstance(n, ast.Name):
                used.add(n.id)
    exported = set()
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "__all__" for t in node.targets):
            if isinstance(node.value, (ast.List, ast.Tuple)):
                exported = {e.value for e in node.value.elts if isinstance(e, ast.Constant)}
    if os.path.basename(path) != "__init__.py":
        for name, lineno in imported.items():
            if name not in used and name not in exported and lineno not in noqa and name != "annotations":
                out.append((lineno, "unused import %r" % name))
    seen = {}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            if node.name in seen and node.lineno not in noqa:
                out.append((node.lineno, "duplicate definition of %r (first at line %d)" % (node.name, seen[node.name])))
            seen[node.name] = node.lineno
    return sorted(out)


def main(argv):
    paths = argv or DEFAULT
    n = 0
    for f in sorted(set(py_files(paths))):
        for lineno, msg in check(f):
            print("%s:%d: %s" % (os.path.relpath(f, ROOT), lineno, msg))
            n += 1
    print("%d finding(s)" % n)
    return 1 if n else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
