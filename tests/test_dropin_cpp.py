"""The C++ drop-in (bsh::InsertionSiteSearch): built as a standalone program against
libbiorseo_b200.so exactly like a patched MOIP.cpp would be, then

  * without a GPU it must fail loudly (there is no CPU fallback behind the boundary),
  * on a GPU its vector<Motif> must equal the reference's insertion_sites_ (multiset).
"""
import os
import subprocess

import pytest

from biorseo_b200 import build, synth

import cases
import util

ROOT = util.ROOT
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "dropin")


def build_dropin():
    """Compiles tests/cpp/dropin_main.cpp against libbiorseo_b200.so (as a patched MOIP.cpp would be) when it is stale."""
    src = os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(build.LIB_PATH)):
        lib_dir = os.path.dirname(build.LIB_PATH)
        subprocess.run(["g++", "-std=c++17", "-O1", src, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "biorseo_b200", "host"),
                        "-L", lib_dir, "-lbiorseo_b200", f"-Wl,-rpath,{lib_dir}", "-o", BIN], check=True)
    return BIN


@pytest.fixture(scope="module")
def dropin():
    return build_dropin()


def _run(binary, source, library, seq, mask, tmp, *extra):
    seq_path, mask_path = util.write_query(str(tmp), seq, mask)
    return subprocess.run([binary, source, str(library), seq_path, mask_path, *extra], capture_output=True, text=True, timeout=300)


def test_fails_loudly_without_a_gpu(dropin, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    case = cases.CASES[0]
    lib = cases.materialise(case, tmp_path)
    res = _run(dropin, case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]), tmp_path)
    assert res.returncode == 3 and res.stdout == ""
    assert "no CUDA device" in res.stderr or "CUDA" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_golden_cases_through_the_cpp_dropin(dropin, case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    res = _run(dropin, case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]), tmp_path)
    assert res.returncode == 0, res.stderr
    assert sorted(line for line in res.stdout.split("\n") if line) == case["expected"]


@pytest.mark.gpu
@pytest.mark.parametrize("source,n", [("jsonfolder", 640), ("rinfolder", 230), ("descfolder", 150)])
def test_random_libraries_through_the_cpp_dropin(dropin, source, n, tmp_path):
    seq = synth.sequence(n, 31)
    mask = synth.mask_canonical(seq, seed=5, keep=0.6)
    if source == "jsonfolder":
        lib = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(200, 9, "c1"))
    elif source == "rinfolder":
        lib = synth.write_rin_folder(str(tmp_path / "rin"), 120, 9, max_components=3, min_len=2)
    else:
        lib = synth.write_desc_folder(str(tmp_path / "desc"), 120, 9)
    res = _run(dropin, source, lib, seq, mask, tmp_path)
    assert res.returncode == 0, res.stderr
    assert sorted(line for line in res.stdout.split("\n") if line) == util.best_oracle(source, lib, seq, mask)


@pytest.mark.gpu
@pytest.mark.parametrize("source,n", [("jsonfolder", 640), ("rinfolder", 230), ("descfolder", 150)])
def test_sharded_dropin_gives_the_same_list(dropin, source, n, tmp_path):
    """bsh::InsertionSiteSearch::open_sharded (one process, the library cut over the GPUs of the box; one range on a
    one-GPU box) against the one-GPU drop-in: the same sites in the same order."""
    import torch
    gpus = min(torch.cuda.device_count(), 4)
    seq = synth.sequence(n, 32)
    mask = synth.mask_canonical(seq, seed=6, keep=0.6)
    if source == "jsonfolder":
        lib = synth.write_json(str(tmp_path / "lib.json"), synth.json_modules(300, 10, "c1"))
    elif source == "rinfolder":
        lib = synth.write_rin_folder(str(tmp_path / "rin"), 120, 10, max_components=3, min_len=2)
    else:
        lib = synth.write_desc_folder(str(tmp_path / "desc"), 120, 10)
    one = _run(dropin, source, lib, seq, mask, tmp_path)
    many = _run(dropin, source, lib, seq, mask, tmp_path, f"gpus={gpus}")
    assert one.returncode == 0 and many.returncode == 0, (one.stderr, many.stderr)
    assert one.stdout and many.stdout == one.stdout


def _csv_lines(rng, n, count):
    lines = ["motif,score,positions"]
    for i in range(count):
        comps = int(rng.integers(1, 4))
        at = int(rng.integers(0, n // 2))
        fields = []
        for c in range(comps):
            k = int(rng.integers(2, 8))
            fields += [at, at + k - 1] if rng.random() > 0.1 or c == 0 else [at + 3, at]     # start >= end: that component is dropped
            at += k + int(rng.integers(1, 25))
        lines.append(f"bp_{i}.{int(rng.integers(1, 99))},{int(rng.integers(-5, 60))}," + ",".join(str(x) for x in fields))
    return lines


@pytest.mark.ref
@pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref is not built")
@pytest.mark.parametrize("seed,n", [(1, 73), (2, 120), (3, 230)])
def test_pre_placed_csv_matches_the_reference(dropin, seed, n, tmp_path):
    """Row "next" #4: the --pre-placed CSV source (cppsrc/MOIP.cpp:114-145). Host-only, so it runs here."""
    import numpy as np
    rng = np.random.default_rng(seed)
    seq = synth.sequence(n, seed)
    mask = synth.mask_canonical(seq, seed=seed, keep=0.9)
    csv = tmp_path / "placed.csv"
    csv.write_text("\n".join(_csv_lines(rng, n, 400)) + "\n")
    res = _run(dropin, "csvfile", csv, seq, mask, tmp_path)
    assert res.returncode == 0, res.stderr
    mine = sorted(line for line in res.stdout.split("\n") if line)
    assert len(mine) > 0
    assert mine == util.run_ref("csvfile", csv, seq, mask, mode="ctor")


def test_pre_placed_csv_skips_lines_the_reference_cannot_read(dropin, tmp_path):
    seq = synth.sequence(60, 4)
    csv = tmp_path / "bad.csv"
    csv.write_text("motif,score,positions\n"
                   "ok,3,2,5,40,44\n"            # kept if the closing pairs are allowed
                   "jar3d,True,12,2,5,-,-\n"     # the reference calls stoi(\"True\")
                   "nonnum,x,2,5\n"
                   "nocomp,3,9,2\n"              # start >= end: no component left, the reference reads comp[0] of an empty vector
                   "short\n")
    res = _run(dropin, "csvfile", csv, seq, synth.mask_all(60), tmp_path)
    assert res.returncode == 0, res.stderr
    assert [line.split("\t")[0] for line in res.stdout.split("\n") if line] == ["ok"]
    assert "4 lines skipped" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A2_json", "A4_rin_canonical", "A5_desc", "seed12_rin"])
def test_variable_index_through_the_cpp_dropin(dropin, name, tmp_path):
    """bsh::InsertionSiteSearch::run(..., &index): the C_{x,i,p} table and the c4 rows the patched constructor
    would use (cppsrc/MOIP.cpp:302-328, 569-586) against the plain-Python restatement."""
    import index_ref
    case = next(c for c in cases.CASES if c["name"] == name)
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    res = _run(dropin, case["source"], lib, seq, mask, tmp_path, "index")
    assert res.returncode == 0, res.stderr
    lines = [line for line in res.stdout.split("\n") if line]
    sites = [line for line in lines if line[:2] not in ("V ", "O ")]
    want = index_ref.build(sites, len(seq), mask, 1 if case["source"] == "descfolder" else 0)
    got_v = [tuple(int(x) for x in line.split()[1:]) for line in lines if line.startswith("V ")]
    exp_v = []
    for x in range(len(sites)):
        for v in range(int(want["var_begin"][x]), int(want["var_begin"][x + 1])):
            exp_v.append((x, v - int(want["var_begin"][x]), int(want["var_first"][v]), int(want["var_second"][v].astype("int32")), int(want["var_c3_terms"][v])))
    assert got_v == exp_v
    got_o = [tuple(int(x) for x in line.split()[1:]) for line in lines if line.startswith("O ")]
    ob, ov = want["overlap_begin"], want["overlap_vars"]
    exp_o = [(u, *[int(v) for v in ov[int(ob[u]):int(ob[u + 1])]]) for u in range(len(seq)) if int(ob[u + 1]) - int(ob[u]) > 1]
    assert got_o == exp_o


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["A2_json", "A4_rin_renumber", "A4_rin_canonical", "A5_desc", "seed11_rin", "seed12_json"])
def test_constraints_through_the_cpp_dropin(dropin, name, tmp_path):
    """The C++ consumer of the variable index (tests/cpp/dropin_main.cpp, `constraints`): c3 / c4 / c5 / c6 built from the
    index instead of the reference's loops must be the reference's own constraints (committed fixture, written from
    `biorseo_ref model`; and the live reference where it is present)."""
    import json
    import index_ref
    case = next(c for c in cases.CASES if c["name"] == name)
    lib = cases.materialise(case, tmp_path)
    seq, mask = case["seq"], cases.mask_of(case["mask"], case["seq"])
    res = _run(dropin, case["source"], lib, seq, mask, tmp_path, "constraints")
    assert res.returncode == 0, res.stderr
    sites, mine = index_ref.parse_model_dump(res.stdout)
    with open(os.path.join(util.GOLDEN, "site_index.json")) as f:
        entry = next(e for e in json.load(f) if e["name"] == name)
    assert sorted(sites) == entry["sites"]
    assert mine == sorted((op, rhs, tuple((n_, c_) for n_, c_ in terms)) for op, rhs, terms in entry["constraints"])
    if util.have_ref():
        import test_site_index
        assert mine == test_site_index.run_model(case["source"], lib, seq, mask)[1]
