"""Poor man's pyflakes (none is installed in the image): reports names that are read somewhere in a file but bound nowhere in
it (module, any function, comprehension, import, argument ...) and are not builtins.  Scope-insensitive on purpose: it only
has to catch typos and leftovers of edits in files whose hot paths cannot be executed without a GPU.
Usage: python tools/undefined_names.py FILE..."""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read(), path)
    bound, loads = set(dir(builtins)) | {"__file__", "__name__"}, []
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            if isinstance(node.ctx, ast.Load):
                loads.append(node)
            else:
                bound.add(node.id)
        elif isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            bound.add(node.name)
        elif isinstance(node, ast.arg):
            bound.add(node.arg)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                bound.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.ExceptHandler) and node.name:
            bound.add(node.name)
        elif isinstance(node, (ast.Global, ast.Nonlocal)):
            bound.update(node.names)
    bad = sorted({(n.id, n.lineno) for n in loads if n.id not in bound})
    for name, line in bad:
        print(f"{path}:{line}: undefined name {name!r}")
    return len(bad)


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
