"""SASS opcode histogram per kernel of libfrozenbilm_sm100.so (runs on the CPU: cuobjdump only).

    python scripts/sass_histogram.py > profiles/rN_sass_histogram.md

For every kernel: total instructions and the counts of the opcodes that prove what the kernel is built from --
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus HMMA / IMMA (legacy mma.sync: expected 0)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "frozenbilm_b200", "libfrozenbilm_sm100.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "IMMA",
       "FFMA", "MUFU", "STS", "LDS", "LDG", "STG", "REDG", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print("# SASS opcode histogram per kernel (cuobjdump -sass frozenbilm_b200/libfrozenbilm_sm100.so)\n")
    print("| kernel | instr | " + " | ".join(KEY) + " |")
    print("|---|---:|" + "---:|" * len(KEY))
    tot = collections.Counter()
    for k, c in kernels.items():
        nm = names[k].replace("void ", "")
        nm = nm[:nm.rfind(">(") + 1] if ">(" in nm else nm.split("(")[0]
        nm = nm.replace("(bool)", "").replace("(int)", "")
        print(f"| `{nm}` | {sum(c.values())} | " + " | ".join(str(c.get(x, 0)) for x in KEY) + " |")
        tot.update(c)
    print(f"| **all {len(kernels)} kernels** | {sum(tot.values())} | " + " | ".join(str(tot.get(x, 0)) for x in KEY) + " |")


if __name__ == "__main__":
    sys.exit(main())
