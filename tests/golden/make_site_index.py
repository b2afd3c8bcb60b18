"""Writes tests/golden/site_index.json: for a few golden libraries, the constraints of the reference's own
constraint builder that involve an insertion variable (cppsrc/MOIP.cpp:302-328, 507-694), produced by
running the reference unmodified against the recording solver shim (oracle/_ref `model` mode) in the build
container. The fixture travels to boxes that have no reference.

    python tests/golden/make_site_index.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import test_site_index as t  # noqa: E402

NAMES = ("A2_json", "A4_rin_renumber", "A4_rin_canonical", "A5_desc", "seed11_rin", "seed12_json")

out = []
for case in cases.CASES:
    if case["name"] not in NAMES:
        continue
    with tempfile.TemporaryDirectory() as tmp:
        lib = cases.materialise(case, tmp)
        sites, constraints = t.run_model(case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]))
    out.append({"name": case["name"], "sites": sorted(sites), "constraints": constraints})
    print(case["name"], len(sites), "sites,", len(constraints), "constraints")
with open(os.path.join(HERE, "site_index.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
