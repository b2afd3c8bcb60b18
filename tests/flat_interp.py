"""Pure-Python interpreter of the flat component table (include/biorseo_b200.h, bss_table).

Test helper only: it lets the CPU test-suite check the host loaders (JSON / RIN / DESC ->
table) against the reference oracle without a GPU, and documents the placement semantics
the kernels implement. Small inputs only. Never used by the product.
"""
import numpy as np

NO_GATE = 0xFFFFFFFF
ABSOLUTE = -1


def _matches(seq, pat, n):
    """All positions q in [0, n] where `pat` ('.' = any byte) matches seq[q:q+len(pat)]."""
    k = len(pat)
    out = []
    for q in range(0, n - k + 1):
        if all(p == 0x2E or p == seq[q + j] for j, p in enumerate(pat)):
            out.append(q)
    return out


def _chain(matches, k, start):
    """Leftmost non-overlapping matches at or after `start` (what sregex_iterator yields)."""
    out, t = [], start
    for q in matches:
        if q >= t:
            out.append(q)
            t = q + max(k, 1)
    return out


def _unit(table, u):
    c0, c1 = int(table["unit_comp_begin"][u]), int(table["unit_comp_begin"][u + 1])
    pats = [bytes(table["pattern"][int(table["comp_pat_begin"][c]):int(table["comp_pat_begin"][c + 1])]) for c in range(c0, c1)]
    k0, k1 = int(table["unit_check_begin"][u]), int(table["unit_check_begin"][u + 1])
    return pats, [tuple(int(x) for x in table["checks"][i]) for i in range(k0, k1)], int(table["unit_delay"][u])


def _allowed(mask, n, u, v):
    a, b = min(u, v), max(u, v)
    if a < 0 or b >= n or a == b:
        return False
    return bool((int(mask[a, b >> 5]) >> (b & 31)) & 1)


def _placements(seq, n, pats, delay):
    per_comp = [_matches(seq, p, n) for p in pats]

    def rec(c, start):
        k = len(pats[c])
        for q in _chain(per_comp[c], k, start):
            if c == len(pats) - 1:
                yield (q,)
                continue
            t = q + k - 1 + delay
            if t < 0 or t >= n:
                break
            for rest in rec(c + 1, t):
                yield (q,) + rest

    return rec(0, 0)


def search_table(table, seq, mask):
    """Returns the u32 record stream [module, p_0..p_{C-1}]* in canonical order."""
    seq = seq.encode() if isinstance(seq, str) else bytes(seq)
    n = len(seq)
    mask = np.asarray(mask, dtype=np.uint32).reshape(n, (n + 31) // 32) if n else np.zeros((0, 0), np.uint32)
    out = []
    for u in range(int(table["n_units"])):
        gate = int(table["unit_gate"][u])
        if gate != NO_GATE:
            gpats, _, gdelay = _unit(table, gate)
            if next(iter(_placements(seq, n, gpats, gdelay)), None) is None:
                continue
        pats, checks, delay = _unit(table, u)
        module = int(table["unit_module"][u])
        for p in _placements(seq, n, pats, delay):
            ok = True
            for (aa, ao, ba, bo) in checks:
                x = (0 if aa == ABSOLUTE else p[aa]) + ao
                y = (0 if ba == ABSOLUTE else p[ba]) + bo
                if not _allowed(mask, n, x, y):
                    ok = False
                    break
            if ok:
                out.extend((module,) + p)
    return np.array(out, dtype=np.uint32)
