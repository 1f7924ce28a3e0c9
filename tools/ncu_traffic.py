#!/usr/bin/env python
"""profiles/r02_traffic.json from an `ncu --set full` report of bench.py: dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per stage,
stamped with the hash of the kernel sources (bench.kernel_sources_sha16) so that bench.py never quotes it for other code.
usage (on the GPU box): tools/ncu_traffic.py /tmp/prof.ncu-rep gpurun_out/r02_traffic.json"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
STAGE = [("me_picture_kernel<1>", "me_subpel"), ("me_picture_kernel<0>", "me_int"), ("ois_", "ois"), ("mc_batch", "mc"), ("tq_batch", "tq"), ("dlf_pass", "dlf"),
         ("sao_stats", "sao_stats"), ("sao_apply", "sao_apply")]


def main():
    import bench
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    acc = {}
    for r in rows[2:]:
        name = r[ik]
        key = next((k for pat, k in STAGE if pat in name.replace("(bool)", "")), None)
        if key is None:
            continue
        b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
        acc.setdefault(key, []).append(b)
    avg = {k: sum(v) / len(v) for k, v in acc.items()}
    per = {k: v for k, v in avg.items() if not k.startswith("me_")}
    if "me_int" in avg:      # the mini-GOP mix of bench.py: seven integer-pel launches, one sub-pel launch
        per["me"] = (7 * avg["me_int"] + avg.get("me_subpel", avg["me_int"])) / 8
    if "dlf" in avg:
        per["dlf"] = 2 * avg["dlf"]      # two passes per picture
    json.dump({"sources_sha16": bench.kernel_sources_sha16(), "report": os.path.basename(rep), "launches": {k: len(v) for k, v in acc.items()},
               "bytes_per_launch": per, "per_kernel": avg}, open(out, "w"), indent=1)
    print(json.dumps(per))


if __name__ == "__main__":
    main()
