"""Lint the python sources of the repository (parity: tools/linter.py, which shells out to autopep8).

Uses autopep8 when installed; otherwise falls back to a built-in check (syntax via ``compile``, line length,
trailing whitespace, tabs) so the command is useful on an air-gapped box."""
import os
import pathlib
import subprocess
import sys

MAX_LINE = 120
CHECK_DIRS = ["megatron_llm_b200", "tools", "tasks", "tests", "weights_conversion", "baseline"]


def python_files():
    root = pathlib.Path(__file__).resolve().parent.parent
    files = set(root.glob("*.py"))
    for sub in CHECK_DIRS:
        files.update((root / sub).rglob("*.py"))
    return sorted(f for f in files if "_build" not in f.parts and "_ref" not in f.parts)


def builtin_check(path):
    problems = []
    src = path.read_text()
    try:
        compile(src, str(path), "exec")
    except SyntaxError as e:
        problems.append(f"{path}:{e.lineno}: syntax error: {e.msg}")
    for i, line in enumerate(src.splitlines(), 1):
        if len(line) > MAX_LINE:
            problems.append(f"{path}:{i}: line longer than {MAX_LINE}")
        if line.rstrip() != line:
            problems.append(f"{path}:{i}: trailing whitespace")
        if "\t" in line:
            problems.append(f"{path}:{i}: tab character")
    return problems


def recursively_lint_files(fix=False):
    files = python_files()
    try:
        import autopep8  # noqa: F401
        have = True
    except ModuleNotFoundError:
        have = False
    if have and fix:
        for f in files:
            print(f)
            subprocess.check_call([sys.executable, "-m", "autopep8", "--max-line-length", str(MAX_LINE),
                                   "--aggressive", "--in-place", str(f)])
        return 0
    problems = [p for f in files for p in builtin_check(f)]
    print("\n".join(problems) if problems else f"{len(files)} files clean")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(recursively_lint_files(fix="--fix" in sys.argv))
