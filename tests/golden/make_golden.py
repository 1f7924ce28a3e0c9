#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.sh where
/root/reference exists).  usage: python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "svt-hevc_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import common  # noqa: E402
import golden_cases  # noqa: E402

if common.load_ref() is None:
    sys.exit("oracle/_ref missing: run oracle/build_ref.sh first")
out = {}
for key, runner, case in golden_cases.cases():
    out[key] = runner(case, "reference")
    print(key, out[key][:16])
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump({"generator": "tests/golden/make_golden.py (reference C rows through oracle/_ref)", "sha256": out}, f, indent=1, sort_keys=True)
