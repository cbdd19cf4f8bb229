"""Per-kernel SASS evidence of the built library (run where cuobjdump is: the build container).
For every kernel of aswstereomatch_b200/libasw_b200.so: instruction count and the mnemonics that matter for the
B200 design claims -- UBLKCP / UTMALDG (TMA bulk copies), LDGSTS (cp.async), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2
(packed FP32), MUFU (SFU), DFMA / DADD / DMUL (FP64), LDL / STL (register spills), BAR.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aswstereomatch_b200", "libasw_b200.so")
KEYS = ["UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "DFMA", "DADD", "DMUL", "LDS", "STS",
        "LDL", "STL", "BAR", "ATOMG", "REDG", "ATOMS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, total = None, {}, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter(); total[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for k in KEYS:
                if op == k or (k in ("BAR",) and op.startswith("BAR")):
                    counts[cur][k] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels (sm_100a)")
    print("# columns: instructions | " + " ".join(KEYS))
    for mangled, name in sorted(zip(counts, demangled), key=lambda t: t[1]):
        short = re.sub(r"\(.*", "", name)
        c = counts[mangled]
        print(f"{short[:70]:70s} {total[mangled]:6d} | " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    main()
