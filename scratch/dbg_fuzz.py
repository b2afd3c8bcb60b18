import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import biorseo_b200 as b
import flat_interp
import test_gpu_parity as t

def run(seed):
    rng = np.random.default_rng(900 + seed)
    alphabet = [ord(c) for c in ("AG" if seed % 3 == 0 else "ACG" if seed % 3 == 1 else "ACGU")]
    n = int([37, 64, 97, 130, 200, 33][seed % 6])
    table = t._random_table(rng, 60, alphabet, max_comp=4 if n < 100 else 3, max_len=4)
    body = [alphabet[i] for i in rng.integers(0, len(alphabet), n)]
    for _ in range(4):
        at, run = int(rng.integers(0, n)), int(rng.integers(3, 12))
        unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
        for j in range(run):
            if at + j < n:
                body[at + j] = unit[j % len(unit)]
    seq = bytes(body)
    span = [n, 12, 30, 5][seed % 4]
    mask = t._banded_random_mask(n, seed, span, [1.0, 0.6, 0.3][seed % 3])
    dev = b.DeviceLibrary.from_table(table)
    got = dev.search(seq, mask).records()
    expect = flat_interp.search_table(table, seq, mask)
    if np.array_equal(got, expect):
        print(seed, "ok"); return
    ncomp = np.diff(table["unit_comp_begin"])
    def split(rec):
        out = {}
        at = 0
        while at < rec.size:
            m = int(rec[at]); w = 1 + int(ncomp[m])
            out.setdefault(m, []).append(tuple(int(x) for x in rec[at+1:at+w])); at += w
        return out
    g, e = split(got), split(expect)
    for m in sorted(set(g) | set(e)):
        if g.get(m) != e.get(m):
            c0, c1 = table["unit_comp_begin"][m], table["unit_comp_begin"][m+1]
            pats = [bytes(table["pattern"][table["comp_pat_begin"][c]:table["comp_pat_begin"][c+1]]) for c in range(c0, c1)]
            k0, k1 = table["unit_check_begin"][m], table["unit_check_begin"][m+1]
            print("unit", m, pats, "delay", table["unit_delay"][m], "checks", table["checks"][k0:k1].tolist())
            gs, es = set(g.get(m, [])), set(e.get(m, []))
            print(" extra", sorted(gs - es)[:10], "missing", sorted(es - gs)[:10], len(g.get(m, [])), len(e.get(m, [])))
            print(" seq", seq.decode())
            break

for s in (int(x) for x in sys.argv[1:]):
    run(s)
