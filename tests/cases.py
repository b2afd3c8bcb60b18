"""Access to the committed golden fixtures (tests/golden/cases.json, produced by the reference)."""
import json
import os

from biorseo_b200 import synth

import util

with open(os.path.join(util.GOLDEN, "cases.json")) as f:
    GOLDEN = json.load(f)

CASES = GOLDEN["cases"]
FNO = GOLDEN["fno"]


def mask_of(spec, seq):
    n = len(seq)
    if spec == "all":
        return synth.mask_all(n)
    if spec == "canonical":
        return synth.mask_canonical(seq)
    _, seed, keep = spec.split(":")
    return synth.mask_canonical(seq, seed=int(seed), keep=float(keep))


def materialise(case, tmp):
    for rel, content in case["files"].items():
        full = os.path.join(str(tmp), "lib", rel)
        os.makedirs(os.path.dirname(full), exist_ok=True)
        with open(full, "w") as f:
            f.write(content)
    if case["source"] == "jsonfolder":
        return os.path.join(str(tmp), "lib", next(iter(case["files"])))
    return os.path.join(str(tmp), "lib")


def fno_table(comps, delay):
    """A one-unit flat table without checks: every placement of find_next_ones_in is a site."""
    import numpy as np
    pat = b"".join(c.encode() for c in comps)
    begins = np.cumsum([0] + [len(c) for c in comps]).astype(np.uint32)
    return {
        "n_modules": 1, "n_units": 1, "n_gates": 0,
        "unit_module": np.zeros(1, np.uint32), "unit_delay": np.array([delay], np.uint32),
        "unit_gate": np.array([0xFFFFFFFF], np.uint32), "unit_comp_begin": np.array([0, len(comps)], np.uint32),
        "comp_pat_begin": begins, "pattern": np.frombuffer(pat, dtype=np.uint8).copy() if pat else np.zeros(0, np.uint8),
        "unit_check_begin": np.zeros(2, np.uint32), "checks": np.zeros((0, 4), np.int16),
    }


def fno_expected_records(entry):
    rec = []
    for line in entry["expected"]:
        rec.append(0)
        rec.extend(int(part.split("-")[0]) for part in line.split(" "))
    return rec
