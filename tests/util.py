"""Shared helpers of the test-suite: run the oracles, write inputs, compare site lists."""
import os
import resource
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "biorseo_ref")      # the reference's own code
PORT_BIN = os.path.join(ROOT, "oracle", "_build", "biorseo_port")  # plain C++ restatement
GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref():
    return os.access(REF_BIN, os.X_OK)


def have_port():
    return os.access(PORT_BIN, os.X_OK)


def write_query(folder, seq, mask):
    seq_path = os.path.join(folder, "seq.txt")
    mask_path = os.path.join(folder, "mask.bin")
    with open(seq_path, "wb") as f:
        f.write(seq.encode() if isinstance(seq, str) else bytes(seq))
        f.write(b"\n")
    np.ascontiguousarray(mask, dtype="<u4").tofile(mask_path)
    return seq_path, mask_path


def _bounded():
    # The reference materialises every placement before filtering it; keep a runaway case
    # from taking the box down.
    resource.setrlimit(resource.RLIMIT_AS, (6 << 30, 6 << 30))


def run_oracle(binary, mode, source, library, seq, mask, threads=2, timeout=120):
    """Runs an oracle CLI; returns its dump as a sorted list of lines (a multiset of sites)."""
    with tempfile.TemporaryDirectory() as tmp:
        seq_path, mask_path = write_query(tmp, seq, mask)
        cmd = [binary, mode, source, str(library), seq_path, mask_path]
        if mode == "jobs":
            cmd.append(str(threads))
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, preexec_fn=_bounded)
    if res.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed ({res.returncode}): {res.stderr[-2000:]}")
    return sorted(line for line in res.stdout.split("\n") if line)


def run_ref(source, library, seq, mask, mode="jobs"):
    return run_oracle(REF_BIN, mode, source, library, seq, mask)


def run_port(source, library, seq, mask):
    return run_oracle(PORT_BIN, "jobs", source, library, seq, mask)


_announced = set()


def best_oracle(source, library, seq, mask, timeout=120, threads=2):
    """The reference's own code (oracle/_ref/biorseo_ref, built from /root/reference by oracle/Makefile and shipped
    to the GPU box prebuilt). Its absence is an error, not a silent switch to the restatement: set BSS_ALLOW_PORT=1
    to accept oracle/port instead (a box where the reference could not be built). Which one ran is printed once."""
    if have_ref():
        which, binary = "reference (oracle/_ref/biorseo_ref)", REF_BIN
    elif os.environ.get("BSS_ALLOW_PORT") == "1" and have_port():
        which, binary = "PORT (oracle/_build/biorseo_port; BSS_ALLOW_PORT=1)", PORT_BIN
    else:
        raise RuntimeError(f"{REF_BIN} is missing: build it with `make -C oracle ref` where /root/reference exists "
                           "(it travels to the GPU box with the snapshot), or set BSS_ALLOW_PORT=1 to check against the C++ restatement instead")
    if which not in _announced:
        _announced.add(which)
        print(f"\n[oracle] parity is checked against the {which}")
    return run_oracle(binary, "jobs", source, library, seq, mask, threads=threads, timeout=timeout)


def ref_placements(source, library, seq, mask):
    """Number of placements the reference's find_next_ones_in returns over a JSON library (before the pair filters)."""
    with tempfile.TemporaryDirectory() as tmp:
        seq_path, mask_path = write_query(tmp, seq, mask)
        res = subprocess.run([REF_BIN, "placements", source, str(library), seq_path, mask_path], capture_output=True, text=True, timeout=300,
                             preexec_fn=_bounded, check=True)
    return int(res.stdout.strip().splitlines()[-1])
