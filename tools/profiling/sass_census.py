"""SASS mnemonic census of the built extension (no GPU needed): which instruction classes every kernel uses.

  python tools/profiling/sass_census.py [path/to/_C_b200.so] > profiles/sass_mnemonics.txt

UTCHMMA / UTCQMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG =
TMA tensor load / store / reduce, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UCGABAR = cluster barrier; HMMA would be
the legacy mma.sync path (expected: absent from the GEMM and attention kernels).
"""
import collections
import re
import subprocess
import sys

INTERESTING = re.compile(r"^(UTC|UTMA|UBLKCP|SYNCS|LDTM|STTM|HMMA|IMMA|QMMA|MUFU|REDG|RED|ATOM|MEMBAR|ERRBAR|FENCE|UCGABAR|"
                         r"ACQBULK|CCTL)")
LINE = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")


def main(path):
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    fn = None
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
            continue
        m = LINE.match(line)
        if m and fn and INTERESTING.match(m.group(1)):
            counts[fn][m.group(1)] += 1
    demangled = {}
    try:
        out = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True, check=True).stdout
        demangled = dict(zip(counts, out.splitlines()))
    except Exception:  # noqa: BLE001 - c++filt is optional
        pass
    for fn in sorted(counts, key=lambda f: demangled.get(f, f)):
        name = demangled.get(fn, fn)
        name = re.sub(r"\(.*", "", name)
        print(name)
        print("    " + "  ".join(f"{k}:{v}" for k, v in sorted(counts[fn].items())))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "megatron_llm_b200/_C_b200.so")
