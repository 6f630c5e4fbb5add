"""Static checks that stand in for a GPU on code paths the CPU suite cannot execute: every Python file parses, and no
function reads a name that nothing in the file binds (imports, definitions, assignments, arguments) - the NameError
class of bugs in CUDA-only branches."""
import ast
import builtins
import glob
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
SKIP = (os.sep + "baseline" + os.sep, os.sep + "_build" + os.sep, "gpurun_out", os.sep + ".git" + os.sep)


def _files():
    for f in glob.glob(os.path.join(ROOT, "**", "*.py"), recursive=True):
        if not any(s in f for s in SKIP):
            yield f


def test_every_python_file_parses_and_binds_the_names_it_reads():
    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__package__"}
    problems = {}
    for path in _files():
        with open(path, errors="ignore") as f:
            tree = ast.parse(f.read(), filename=path)          # SyntaxError fails the test with the location
        bound, star = set(module_names), False
        for n in ast.walk(tree):
            if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                bound.add(n.name)
            elif isinstance(n, ast.Import):
                bound.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, ast.ImportFrom):
                star = star or any(a.name == "*" for a in n.names)
                bound.update(a.asname or a.name for a in n.names)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                bound.add(n.id)
            elif isinstance(n, ast.arg):
                bound.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                bound.add(n.name)
            elif isinstance(n, (ast.Global, ast.Nonlocal)):
                bound.update(n.names)
            elif isinstance(n, ast.MatchAs) and n.name:
                bound.add(n.name)
        if star:
            continue
        read = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
        missing = sorted(read - bound)
        if missing:
            problems[os.path.relpath(path, ROOT)] = missing
    assert not problems, problems
