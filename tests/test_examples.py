"""The launch scripts under examples/ must stay runnable: valid shell syntax, and every ``--flag`` they pass must be a
flag some parser in this repository (or the launcher) defines.  (The reference's examples reference flags that no
longer exist; this keeps ours from rotting the same way.)"""
import glob
import os
import re
import subprocess

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), os.pardir))
SCRIPTS = sorted(glob.glob(os.path.join(ROOT, "examples", "**", "*.sh"), recursive=True))
# flags of torchrun / srun / sbatch / helper scripts' own option parsing
LAUNCHER = {"--nproc_per_node", "--nproc-per-node", "--nnodes", "--node_rank", "--node-rank", "--master_addr",
            "--master-addr", "--master_port", "--master-port", "--standalone", "--local-addr", "--rdzv_backend",
            "--rdzv_endpoint", "--rdzv-backend", "--rdzv-endpoint", "--help", "--ntasks-per-node", "--container-image",
            "--container-mounts", "--export", "--nodes", "--job-name", "--time", "--partition", "--account",
            "--gpus-per-node", "--output", "--error", "--exclusive", "--mem", "--cpus-per-task", "--no-container-mount-home"}


def _defined_flags():
    flags = set()
    pat = re.compile(r"""["'](--[A-Za-z][A-Za-z0-9_\-]*)["']""")
    for path in glob.glob(os.path.join(ROOT, "**", "*.py"), recursive=True):
        if any(part in path for part in (os.sep + "baseline" + os.sep, os.sep + "_build" + os.sep, "gpurun_out")):
            continue
        with open(path, errors="ignore") as f:
            flags.update(pat.findall(f.read()))
    # examples/_common.sh and the wrappers parse their own short option set
    for path in SCRIPTS:
        with open(path) as f:
            flags.update(re.findall(r"(--[a-z][a-z0-9_\-]*)\)", f.read()))      # `case` arms: --opt) ...
    return flags


def test_example_scripts_have_valid_syntax():
    assert len(SCRIPTS) >= 20
    for s in SCRIPTS:
        r = subprocess.run(["bash", "-n", s], capture_output=True, text=True)
        assert r.returncode == 0, f"{s}: {r.stderr}"


def test_example_scripts_only_use_known_flags():
    known = _defined_flags() | LAUNCHER
    unknown = {}
    for s in SCRIPTS:
        text = open(s).read()
        text = re.sub(r"#.*", "", text)
        used = set(re.findall(r"(?<![\w-])(--[A-Za-z][A-Za-z0-9_\-]*)", text))
        bad = sorted(f for f in used if f not in known)
        if bad:
            unknown[os.path.relpath(s, ROOT)] = bad
    assert not unknown, unknown
