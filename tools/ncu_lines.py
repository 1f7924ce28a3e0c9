#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel: joins `ncu --page source --csv` (SASS order) with
`nvdisasm -g` line annotations of the same cubin.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-substring>[:<mangled-substring>] <file.cu> [top]
(the second form when the report's demangled name and the cubin's mangled name differ, e.g. template instantiations)"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(kernel, cu):
    so = os.path.join(ROOT, "svt-hevc_b200", "libsvthevc_b200.so")
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
    base = os.path.basename(cu).replace(".cu", "")
    cub = [f for f in os.listdir(d) if f.startswith(base + ".")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
    out, cur, infn = [], None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            infn = kernel in ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    return out


def main():
    rep, kernel, cu = sys.argv[1:4]
    kernel, _, mangled = kernel.partition(":")
    mangled = mangled or kernel
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # one section per captured launch (and per view): "Kernel Name" row, header row, SASS rows
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    lines = sass_lines(mangled, cu)
    def parse(sec):
        hdr = sec["rows"][0]
        ie, st = hdr.index("Instructions Executed"), hdr.index("# Samples")
        return [(int(r[ie] or 0), int(r[st] or 0)) for r in sec["rows"][1:] if len(r) > ie and r[0].startswith("0x")]
    parsed = [(sec["name"], parse(sec)) for sec in secs if sec["rows"] and "Instructions Executed" in sec["rows"][0]]
    # the launch whose SASS length matches the cubin's function (template instantiations share the base name)
    name, data = next(((n, d) for n, d in parsed if len(d) == len(lines)), parsed[0])
    print(f"# launch: {name}")
    if len(lines) != len(data):
        print(f"# warning: {len(data)} SASS rows in the report vs {len(lines)} in the cubin (binary changed since the capture?)")
    n = min(len(lines), len(data))
    agg = {}
    for (inst, smp), key in zip(data[:n], lines[:n]):
        a = agg.setdefault(key, [0, 0])
        a[0] += inst; a[1] += smp
    ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    src = {}
    print(f"# {kernel}: {ti} warp instructions, {ts} stall samples")
    for key, (inst, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            f = os.path.join(ROOT, "svt-hevc_b200", "csrc", key[0])
            if f not in src and os.path.exists(f):
                src[f] = open(f).read().splitlines()
            if f in src and key[1] - 1 < len(src[f]):
                text = src[f][key[1] - 1].strip()[:100]
        print(f"{inst / ti * 100:5.1f}% inst {smp / max(ts, 1) * 100:5.1f}% stall  {key[0] if key else '?'}:{key[1] if key else 0:<5d} {text}")


if __name__ == "__main__":
    main()
