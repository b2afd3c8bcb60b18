"""Generates tests/golden/cases.json by running the REFERENCE's own code (oracle/_ref/biorseo_ref,
built from /root/reference/cppsrc by `make -C oracle ref`) on small hand-written and seeded
inputs. Run in the build container only; the committed JSON travels to the GPU box.

    python tests/golden/make_golden.py

Each case stores the library files, the query, the pair mask (hex of the packed words) and the
reference's dump, sorted. `ctor` cases ran the real MOIP constructor (cppsrc/MOIP.cpp:58); the
`fno` cases ran find_next_ones_in (cppsrc/Motif.cpp:432) directly. The hand-written cases are
the vectors of SURVEY.md appendix A.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import synth  # noqa: E402
import util  # noqa: E402

Q60 = "ACUAGCGAAGGCUAAGUCCGGGAAACCCAAGGAUUUUUUCCAAGGAAAAAAUCCGCAAAAA"
Q48 = "AAGGACCCCCCCUCCGGGGGGGGGGAAAAAUUUUUUUUUUUUUUUU"


def jlib(mods):
    return {"lib.json": json.dumps({k: {"sequence": s, "struct2d": ss} for k, s, ss in mods})}


def rin(links, comps):
    lines = ["ntA,ntB,long_range;...", links, "pos;k;seq"]
    at = 0
    for c in comps:
        lines.append(f"{at},{at + len(c) - 1};{len(c)};{c}")
        at += len(c)
    return "\n".join(lines) + "\n"


def desc(bases, inter=()):
    return "id: 1\nBases: " + "".join(f"{p}_{nt}  " for p, nt in bases) + "\n" + "".join(f"\t({a}) {t} ({b})\n" for a, t, b in inter)


def hand_cases():
    cases = []
    # A.2 JSON end to end
    cases.append(dict(name="A2_json", source="jsonfolder", seq=Q60, mask="canonical",
                      files=jlib([("1", "ACUAGCG&GGCUA&GU", "((((((.&.))))&))"), ("hp", "GGGAAACCC", "(((...)))"),
                                  ("il", "GGA&UCC", "((.&.))"), ("nolink", "GGAA", "...."), ("bad", "GGXA", "(..)")])))
    # A.3 link mapping / alphabet
    cases.append(dict(name="A3_json_alignment", source="jsonfolder", seq=Q48, mask="all",
                      files=jlib([("aligned", "GGA&UCC", "((.&.))"), ("misaligned", "GGA&UCC", "((&..))"),
                                  ("lower_t", "gga&tcc", "((.&.))")])))
    cases.append(dict(name="A3_json_T", source="jsonfolder", seq="AAGGACCCCCCCTCCGGGGGGGGGG", mask="all",
                      files=jlib([("lower_t", "gga&tcc", "((.&.))"), ("upper_u", "GGA&UCC", "((.&.))")])))
    cases.append(dict(name="json_brackets_lowercase_query", source="jsonfolder", seq="GGaGGACGUACGUACGUGGAAACCUU", mask="all",
                      files=jlib([("pk", "GGA&ACG&GGAAACC", "([.&.<)&].....>"), ("sq", "ACGUACG", "[.{.}.]"),
                                  ("case", "GGAGG", "(...)"), ("emptymid", "GG&&ACGU", "((&&..))")])))
    # A.4 RIN renumbering
    cases.append(dict(name="A4_rin_renumber", source="rinfolder", seq=Q48, mask="all",
                      files={"Subfiles/0.txt": rin("0,5,False;1,8,True;3,4,False;4,9,False;", ["GGA", "UCC", "AAAAA"])}))
    cases.append(dict(name="A4_rin_canonical", source="rinfolder", seq=Q60, mask="canonical",
                      files={"Subfiles/7.txt": rin("0,1,False;", ["GC", "AAA"]), "Subfiles/0.txt": rin("0,5,False;1,4,True;", ["GGA", "UCC"]),
                             "Subfiles/3.txt": rin("", ["GGA", "UCC"]), "Subfiles/4.txt": rin("0,1,True;", ["GG", "A"])}))
    # A.5 DESC
    hp = [(10, "G"), (11, "G"), (13, "A"), (14, "A"), (15, "C"), (16, "C")]
    il = [(10, "G"), (11, "G"), (12, "A"), (30, "U"), (31, "C"), (32, "C")]
    short2 = [(10, "G"), (11, "G"), (30, "U"), (31, "C"), (32, "C")]
    short1 = [(10, "G"), (11, "G"), (12, "A"), (30, "U")]
    mid2 = [(10, "G"), (11, "G"), (12, "A"), (30, "U"), (31, "C"), (50, "A"), (51, "A"), (52, "A")]
    mid1 = [(10, "G"), (11, "G"), (12, "A"), (30, "U"), (50, "A"), (51, "A"), (52, "A")]
    last2 = [(10, "G"), (11, "G"), (12, "A"), (30, "U"), (31, "C")]
    bad_nt = [(10, "G"), (11, "N"), (12, "A"), (13, "A")]
    cases.append(dict(name="A5_desc", source="descfolder", seq=Q60, mask="canonical",
                      files={"hp.desc": desc(hp, [("10_G", "+/+", "16_C")]), "il.desc": desc(il, [("10_G", "C/C", "11_G")]),
                             "short2.desc": desc(short2), "short1.desc": desc(short1), "mid2.desc": desc(mid2),
                             "mid1.desc": desc(mid1), "last2.desc": desc(last2), "bad_nt.desc": desc(bad_nt),
                             "bad_cc.desc": desc(il, [("10_G", "C/C", "12_A")]), "bad_hp.desc": desc(il, [("10_G", "+/+", "12_A")])}))
    cases.append(dict(name="desc_all_mask", source="descfolder", seq=Q60, mask="all",
                      files={"hp.desc": desc(hp), "il.desc": desc(il), "short2.desc": desc(short2), "short1.desc": desc(short1),
                             "mid2.desc": desc(mid2), "mid1.desc": desc(mid1), "last2.desc": desc(last2)}))
    return cases


def seeded_cases():
    cases = []
    for seed, n in [(11, 73), (12, 140)]:
        seq = synth.sequence(n, seed)
        with tempfile.TemporaryDirectory() as tmp:
            synth.write_json(os.path.join(tmp, "lib.json"), synth.json_modules(120, seed, "c1"))
            synth.write_json(os.path.join(tmp, "fuzz.json"), [m for m in synth.json_fuzz_modules(150, seed)])
            synth.write_rin_folder(os.path.join(tmp, "rin"), 40, seed)
            synth.write_desc_folder(os.path.join(tmp, "desc"), 40, seed)

            def grab(rel):
                out = {}
                base = os.path.join(tmp, rel)
                if os.path.isfile(base):
                    return {os.path.basename(rel): open(base).read()}
                for dirpath, _, names in os.walk(base):
                    for nm in names:
                        full = os.path.join(dirpath, nm)
                        out[os.path.relpath(full, base)] = open(full).read()
                return out
            cases.append(dict(name=f"seed{seed}_json", source="jsonfolder", seq=seq, mask=f"canonical:{seed}:0.7", files=grab("lib.json")))
            cases.append(dict(name=f"seed{seed}_jsonfuzz", source="jsonfolder", seq=seq, mask="all", files=grab("fuzz.json")))
            cases.append(dict(name=f"seed{seed}_rin", source="rinfolder", seq=seq, mask="all", files=grab("rin")))
            cases.append(dict(name=f"seed{seed}_desc", source="descfolder", seq=seq, mask="all", files=grab("desc")))
    return cases


FNO = [("GGGGG", 1, ["GG"]), ("GGGGGG", 1, ["GG", "GG"]), ("GACGUCGGC", 1, ["G.C"]), ("GGAAAACCACC", 5, ["GG", "CC"]),
       ("GGAAACCACC", 5, ["GG", "CC"]), ("GGaGG", 1, ["GA"]), ("GGaGG", 1, ["G.G"]), ("AGGCAGG", 1, ["GG", "C"]),
       ("ACGUACGUACGU", 1, ["AC", "GU", "AC"]), ("AAAAGGGGGG", 1, ["AA", "GGG"]), ("AAGG", 1, ["", "GG"]), ("AAGG", 1, ["AA", ""]),
       ("AAAAAAAAAA", 1, ["AA", "A.A"]), ("ACACACACACAC", 2, ["ACA", "CAC"]), ("GGGGGGGGGGGG", 0, ["GG", "GG", "GG"])]


def main():
    assert util.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {"generator": "tests/golden/make_golden.py", "reference": "persalteas/biorseo v2.1, unmodified cppsrc", "cases": [], "fno": []}
    for case in hand_cases() + seeded_cases():
        mask = mask_of(case["mask"], case["seq"])
        with tempfile.TemporaryDirectory() as tmp:
            lib = materialise(case, tmp)
            expected = util.run_ref(case["source"], lib, case["seq"], mask, mode="ctor")
            assert expected == util.run_ref(case["source"], lib, case["seq"], mask, mode="jobs"), case["name"]
        case["expected"] = expected
        out["cases"].append(case)
        print(f"{case['name']:32s} {len(expected):5d} sites")
    for seq, delay, comps in FNO:
        res = subprocess.run([util.REF_BIN, "fno", str(delay), seq] + comps, capture_output=True, text=True, check=True)
        out["fno"].append(dict(seq=seq, delay=delay, comps=comps, expected=[line for line in res.stdout.split("\n") if line]))
    with open(os.path.join(util.GOLDEN, "cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote", os.path.join(util.GOLDEN, "cases.json"), os.path.getsize(os.path.join(util.GOLDEN, "cases.json")), "bytes")


def mask_of(spec, seq):
    n = len(seq)
    if spec == "all":
        return synth.mask_all(n)
    if spec == "canonical":
        return synth.mask_canonical(seq)
    kind, seed, keep = spec.split(":")
    return synth.mask_canonical(seq, seed=int(seed), keep=float(keep))


def materialise(case, tmp):
    """Writes the case's library under tmp and returns the path to hand to the loaders."""
    for rel, content in case["files"].items():
        full = os.path.join(tmp, "lib", rel)
        os.makedirs(os.path.dirname(full), exist_ok=True)
        with open(full, "w") as f:
            f.write(content)
    if case["source"] == "jsonfolder":
        return os.path.join(tmp, "lib", next(iter(case["files"])))
    return os.path.join(tmp, "lib")


def make_pair_mask_golden():
    """tests/golden/pair_mask.json: masks produced by the reference's own MOIP constructor
    (oracle/_ref `pairmask` mode) for the seeded matrices of tests/test_pair_mask.py."""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_pair_mask as t
    cases = []
    for n, seed, theta in t.CASES:
        with tempfile.TemporaryDirectory() as tmp:
            ref = t.reference_mask(synth.sequence(n, seed), t.matrix(n, seed, theta), theta, tmp)
        cases.append({"n": n, "seed": seed, "theta": theta, "mask_hex": [f"{int(x):08x}" for x in ref.reshape(-1)]})
    with open(os.path.join(ROOT, "tests", "golden", "pair_mask.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py::make_pair_mask_golden (reference MOIP constructor)", "cases": cases}, f)


if __name__ == "__main__":
    if "--pair-mask" in sys.argv:
        make_pair_mask_golden()
    else:
        main()
