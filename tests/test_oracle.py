"""The oracle is only trusted once pinned: the C++ restatement (oracle/port) must reproduce
what the reference's own code produced for the committed golden vectors, and — where the
compiled reference is available — agree with it on fresh seeded inputs."""
import subprocess

import pytest

from biorseo_b200 import synth

import cases
import util

needs_port = pytest.mark.skipif(not util.have_port(), reason="oracle/_build/biorseo_port not built")
needs_ref = pytest.mark.skipif(not util.have_ref(), reason="oracle/_ref/biorseo_ref not built (needs /root/reference)")


@needs_port
@pytest.mark.parametrize("case", cases.CASES, ids=[c["name"] for c in cases.CASES])
def test_port_reproduces_golden(case, tmp_path):
    lib = cases.materialise(case, tmp_path)
    got = util.run_port(case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]))
    assert got == case["expected"]


@needs_port
@pytest.mark.parametrize("entry", cases.FNO, ids=[f"{e['seq']}:{'+'.join(e['comps'])}" for e in cases.FNO])
def test_port_enumerator_golden(entry):
    res = subprocess.run([util.PORT_BIN, "fno", str(entry["delay"]), entry["seq"]] + entry["comps"], capture_output=True, text=True, check=True)
    assert [line for line in res.stdout.split("\n") if line] == entry["expected"]


@needs_ref
@pytest.mark.parametrize("case", cases.CASES[:8], ids=[c["name"] for c in cases.CASES[:8]])
def test_reference_still_reproduces_golden(case, tmp_path):
    """Guards the fixtures themselves: the real MOIP constructor gives the committed answer."""
    lib = cases.materialise(case, tmp_path)
    got = util.run_ref(case["source"], lib, case["seq"], cases.mask_of(case["mask"], case["seq"]), mode="ctor")
    assert got == case["expected"]


@needs_ref
@needs_port
@pytest.mark.parametrize("seed", range(6))
def test_port_matches_reference_on_random_libraries(seed, tmp_path):
    n = 50 + 37 * seed
    seq = synth.sequence(n, 500 + seed)
    mask = synth.mask_all(n) if seed % 2 else synth.mask_canonical(seq, seed=seed, keep=0.8)
    libs = [("jsonfolder", synth.write_json(str(tmp_path / "a.json"), synth.json_modules(150, seed, "c1"))),
            ("jsonfolder", synth.write_json(str(tmp_path / "f.json"), synth.json_fuzz_modules(150, seed))),
            ("rinfolder", synth.write_rin_folder(str(tmp_path / "rin"), 60, seed)),
            ("descfolder", synth.write_desc_folder(str(tmp_path / "desc"), 60, seed))]
    for source, lib in libs:
        assert util.run_port(source, lib, seq, mask) == util.run_ref(source, lib, seq, mask), (source, seed)
