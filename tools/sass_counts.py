#!/usr/bin/env python3
"""Static SASS opcode counts per kernel of libsvthevc_b200.so (cuobjdump -sass), written to profiles/rNN_sass_counts.txt.

The columns are the ones the review of round 1 asked for: generic loads / stores (LD / ST without a state space: none may remain in a kernel whose pointers
are all known global or shared), local memory (LDL / STL: register spills or dynamically indexed arrays), the global / shared / TMA data movement and the
byte-SAD instruction.  Static counts: they say what the compiler emitted, the dynamic mix comes from the ncu source page (profiles/*_lines_*.txt).
  python tools/sass_counts.py [lib.so] > profiles/r02_sass_counts.txt
"""
import collections
import re
import subprocess
import sys

COLS = ["total", "LD", "ST", "LDL", "STL", "LDG", "STG", "LDS", "STS", "LDSM", "UTMALDG", "UBLKCP", "SYNCS", "VABSDIFF4", "VABSDIFF", "IDP", "IMAD", "IADD3", "LOP3", "SHF", "PRMT",
        "SHFL", "REDUX", "BAR", "ATOMS", "ATOMG", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "svt-hevc_b200/libsvthevc_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, name = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("svtb200::", "").replace("void ", "")
            counts[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            op = m.group(1)
            counts[name]["total"] += 1
            counts[name][op] += 1
    w = max(len(n) for n in counts) + 1
    print("# static SASS opcode counts per kernel (cuobjdump -sass %s); LD/ST = generic address space, LDL/STL = local memory" % lib)
    print("kernel".ljust(w) + " ".join(c.rjust(9) for c in COLS))
    for n in sorted(counts):
        print(n.ljust(w) + " ".join(str(counts[n][c]).rjust(9) for c in COLS))


if __name__ == "__main__":
    main()
