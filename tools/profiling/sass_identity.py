"""Prove on a machine without a GPU that a refactor left a kernel's machine code untouched: compile two versions of a
``.cu`` file to cubins (``nvcc -cubin -gencode arch=compute_100a,code=sm_100a ...``) and compare the SASS of every kernel
of the NEW file whose added template flags are zero against the kernel of the same name and head dim in the OLD file.

    python tools/profiling/sass_identity.py old.cubin new.cubin

Addresses / encodings in comments are ignored; everything else (opcodes, registers, constant-bank offsets, immediates)
must match line by line.  Exit status 1 on any difference.  Used for the attention kernels when the fp16 / dropout
template flags were added (profiles/attention_variants_sass_identity_r2.txt)."""
import difflib
import re
import subprocess
import sys


def functions(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = []
            out[m.group(1)] = cur
            continue
        if cur is not None:
            line = re.sub(r"/\*[0-9a-f]{4,}\*/", "", line).strip()
            if line and not line.startswith((".", "//")):
                cur.append(line)
    return out


def key(name):
    """(kernel name, first integer template argument) of a mangled mlb:: kernel"""
    m = re.match(r"_ZN3mlb\d+(\w+?)ILi(\d+)", name)
    return m.groups() if m else (name, "")


def main(old, new):
    base, cand = functions(old), functions(new)
    by_key = {key(k): v for k, v in base.items()}
    ok = True
    for name, body in cand.items():
        if re.search(r"ILi\d+ELi[1-9]\d*E", name):      # a non-zero flag instantiation: new code, nothing to compare
            continue
        ref = by_key.get(key(name))
        if ref is None:
            print("NEW       ", name)
            continue
        same = ref == body
        ok &= same
        print(("IDENTICAL " if same else "DIFFERENT "), name, f"({len(body)} lines)")
        if not same:
            print("\n".join(list(difflib.unified_diff(ref, body, lineterm="", n=0))[:40]))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
