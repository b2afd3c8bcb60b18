"""Random flat component tables and banded random pair masks shared by the GPU parity tests and the CPU
emulation test of the window walk: chains of overlapping matches x band windows x every check shape."""
import numpy as np

import biorseo_b200 as b


def random_table(rng, n_units, alphabet, max_comp, max_len, min_len=1, short_limit=None):
    """short_limit: units holding a pattern shorter than min_len + 1 get at most that many components
    (keeps the number of placements the Python interpreter has to enumerate bounded)."""
    pats, comp_begin, unit_comp_begin, unit_check_begin, checks, delays = [], [0], [0], [0], [], []
    for _ in range(n_units):
        C = int(rng.integers(1, max_comp + 1))
        lens = []
        for _c in range(C):
            k = int(rng.integers(0 if rng.random() < 0.05 else min_len, max_len + 1))
            if short_limit is not None and k <= min_len and _c >= short_limit:
                k = max_len
            p = [alphabet[i] for i in rng.integers(0, len(alphabet), k)]
            p = [0x2E if rng.random() < 0.15 else x for x in p]
            pats.extend(p)
            comp_begin.append(len(pats))
            lens.append(k)
        unit_comp_begin.append(len(comp_begin) - 1)
        for _k in range(int(rng.integers(0, 4))):
            kind = rng.random()
            a = int(rng.integers(0, C))
            b = int(rng.integers(0, C))
            if kind < 0.15:
                a = -1
            elif kind < 0.25:
                b = a
            ao = int(rng.integers(0, 40)) if a < 0 else int(rng.integers(-2, max(lens[a], 1) + 2))
            bo = int(rng.integers(0, 40)) if b < 0 else int(rng.integers(-2, max(lens[b], 1) + 2))
            checks.append((a, ao, b, bo))
        unit_check_begin.append(len(checks))
        delays.append(int(rng.integers(0, 6)))
    return {
        "n_modules": n_units, "n_units": n_units, "n_gates": 0,
        "unit_module": np.arange(n_units, dtype=np.uint32), "unit_delay": np.array(delays, np.uint32),
        "unit_gate": np.full(n_units, 0xFFFFFFFF, np.uint32), "unit_comp_begin": np.array(unit_comp_begin, np.uint32),
        "comp_pat_begin": np.array(comp_begin, np.uint32), "pattern": np.array(pats, np.uint8) if pats else np.zeros(0, np.uint8),
        "unit_check_begin": np.array(unit_check_begin, np.uint32),
        "checks": np.array(checks, np.int16).reshape(-1, 4),
    }


def banded_random_mask(n, seed, span, density):
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return b.pack_mask((rng.random((n, n)) < density) & (v > u) & (v - u <= span))


def window_table(rng, n_units, alphabet, max_comp, max_len, min_len=1):
    """Units in which every component after the first is tied to an earlier one by a "row check" (an end on an
    earlier component at or before its last nucleotide — possibly before its first —, the other end on the
    component at an offset >= 0, delay >= 1): what JSON / DESC closing pairs look like, plus ends outside their
    component, empty components, extra checks of every other shape, self-overlapping patterns."""
    pats, comp_begin, unit_comp_begin, unit_check_begin, checks, delays = [], [0], [0], [0], [], []
    for _ in range(n_units):
        C = int(rng.integers(1, max_comp + 1))
        lens = []
        for _c in range(C):
            k = int(rng.integers(0 if (rng.random() < 0.08 and min_len <= 1) else min_len, max_len + 1))
            if rng.random() < 0.25 and k >= 2:      # periodic pattern: matches overlap
                unit = [alphabet[i] for i in rng.integers(0, len(alphabet), int(rng.integers(1, 3)))]
                p = [unit[j % len(unit)] for j in range(k)]
            else:
                p = [alphabet[i] for i in rng.integers(0, len(alphabet), k)]
            p = [0x2E if rng.random() < 0.15 else x for x in p]
            pats.extend(p)
            comp_begin.append(len(pats))
            lens.append(k)
        unit_comp_begin.append(len(comp_begin) - 1)
        for c in range(1, C):
            for _r in range(1 if rng.random() < 0.7 else 2):
                a = int(rng.integers(0, c))
                ao = int(rng.integers(-1 if rng.random() < 0.2 else 0, max(lens[a], 1))) if lens[a] else -1
                ao = min(ao, lens[a] - 1)
                bo = int(rng.integers(0, max(lens[c], 1) + (2 if rng.random() < 0.15 else 0)))
                pair = (a, ao, c, bo) if rng.random() < 0.5 else (c, bo, a, ao)    # either order in the table
                checks.append(pair)
        for _k in range(int(rng.integers(0, 3))):    # extra checks of any shape
            kind = rng.random()
            a = int(rng.integers(0, C))
            bb = int(rng.integers(0, C))
            if kind < 0.2:
                a = -1
            elif kind < 0.4:
                bb = a
            ao = int(rng.integers(0, 40)) if a < 0 else int(rng.integers(-2, max(lens[a], 1) + 2))
            bo = int(rng.integers(0, 40)) if bb < 0 else int(rng.integers(-2, max(lens[bb], 1) + 2))
            checks.append((a, ao, bb, bo))
        unit_check_begin.append(len(checks))
        delays.append(int(rng.integers(1, 6)))
    return {
        "n_modules": n_units, "n_units": n_units, "n_gates": 0,
        "unit_module": np.arange(n_units, dtype=np.uint32), "unit_delay": np.array(delays, np.uint32),
        "unit_gate": np.full(n_units, 0xFFFFFFFF, np.uint32), "unit_comp_begin": np.array(unit_comp_begin, np.uint32),
        "comp_pat_begin": np.array(comp_begin, np.uint32), "pattern": np.array(pats, np.uint8) if pats else np.zeros(0, np.uint8),
        "unit_check_begin": np.array(unit_check_begin, np.uint32),
        "checks": np.array(checks, np.int16).reshape(-1, 4),
    }
