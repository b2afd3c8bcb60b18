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


@pytest.fixture(scope="module")
def dropin():
    src = os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(build.LIB_PATH)):
        lib_dir = os.path.dirname(build.LIB_PATH)
        subprocess.run(["g++", "-std=c++17", "-O1", src, "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "biorseo_b200", "host"),
                        "-L", lib_dir, "-lbiorseo_b200", f"-Wl,-rpath,{lib_dir}", "-o", BIN], check=True)
    return BIN


def _run(binary, source, library, seq, mask, tmp):
    seq_path, mask_path = util.write_query(str(tmp), seq, mask)
    return subprocess.run([binary, source, str(library), seq_path, mask_path], capture_output=True, text=True, timeout=300)


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
