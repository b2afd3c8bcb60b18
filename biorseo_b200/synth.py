"""Seeded synthetic inputs in the reference's on-disk formats.

The motif libraries the reference uses are not shipped with it (data/modules is
git-ignored, SOURCES.md:1-69) and ViennaRNA is absent, so benchmarks and tests synthesise
  * query sequences (uniform ACGU unless stated),
  * JSON libraries in the SOURCES.md:14-23 format,
  * CaRNAval RIN files in the format written by scripts/Install_CaRNAval_RINs.py:54-101,
  * Rna3Dmotif .desc files in the format read by cppsrc/Motif.cpp:218-261,
  * pair masks standing in for `pij > theta` (cppsrc/MOIP.cpp:77-90).
Shapes follow SURVEY.md section 8(d).
"""
import json
import os

import numpy as np

from .api import pack_mask

BASES = "ACGU"
CANONICAL = ["AU", "UA", "GC", "CG", "GU", "UG"]


def sequence(n, seed):
    rng = np.random.default_rng(seed)
    return "".join(BASES[i] for i in rng.integers(0, 4, n))


def moip_rules(allowed):
    """Applies the structural rules of MOIP::allowed_basepair (cppsrc/MOIP.cpp:902-904) to a
    boolean n x n matrix: v - u >= 4 and u < n - 6."""
    allowed = np.asarray(allowed, dtype=bool)
    n = allowed.shape[0]
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return allowed & (v - u >= 4) & (u < n - 6)


def mask_canonical(seq, seed=None, keep=0.3, span=100):
    """Canonical pairs, 4 <= v-u < span, u < n-6, thinned by Bernoulli(keep)."""
    s = np.frombuffer(seq.encode(), dtype=np.uint8)
    n = len(s)
    pair = np.zeros((256, 256), dtype=bool)
    for p in CANONICAL:
        pair[ord(p[0]), ord(p[1])] = True
    m = pair[s[:, None], s[None, :]]
    u, v = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    m &= (v - u < span)
    if seed is not None:
        m &= np.random.default_rng(seed).random((n, n)) < keep
    return pack_mask(moip_rules(m))


def mask_all(n):
    return pack_mask(moip_rules(np.ones((n, n), dtype=bool)))


def mask_random(n, seed, density=0.5):
    rng = np.random.default_rng(seed)
    return pack_mask(moip_rules(rng.random((n, n)) < density))


# --------------------------------------------------------------------------------------
# JSON
# --------------------------------------------------------------------------------------
def _strand_struct(lengths):
    """Loop-closing pairs only: first nt of strand 0 with last nt of the last strand, and the
    last nt of each strand with the first nt of the next ('(..(&)..)' style)."""
    out = []
    last = len(lengths) - 1
    for i, k in enumerate(lengths):
        s = ["."] * k
        if len(lengths) == 1:
            s[0], s[-1] = "(", ")"
        else:
            s[0] = "(" if i == 0 else ")"
            s[-1] = ")" if i == last else "("
        out.append("".join(s))
    return "&".join(out)


def _strand_seqs(rng, lengths, pair_bias):
    strands = [[BASES[i] for i in rng.integers(0, 4, k)] for k in lengths]
    # make closing pairs canonical with probability pair_bias
    ends = []
    last = len(lengths) - 1
    if len(lengths) == 1:
        ends.append(((0, 0), (0, lengths[0] - 1)))
    else:
        ends.append(((0, 0), (last, lengths[last] - 1)))
        for i in range(last):
            ends.append(((i, lengths[i] - 1), (i + 1, 0)))
    for (a, b) in ends:
        if rng.random() < pair_bias:
            p = CANONICAL[rng.integers(0, 6)]
            strands[a[0]][a[1]] = p[0]
            strands[b[0]][b[1]] = p[1]
    return "&".join("".join(s) for s in strands)


def json_modules(n_modules, seed, kind="c1", pair_bias=0.8):
    """List of (key, sequence, struct2d). kind: 'c1' mixed hairpin / internal loop / 3-way
    junction; 'c4a' 2-4 strands of 5-8 nt; 'c4b' 3 strands of 3-4 nt; 'mixed' half c4a half c4b."""
    rng = np.random.default_rng(seed)
    mods = []
    for i in range(n_modules):
        k = kind
        if kind == "mixed":
            k = "c4a" if i % 2 == 0 else "c4b"
        if k == "c1":
            r = rng.random()
            if r < 0.40:
                lengths = [int(rng.integers(5, 13))]
            elif r < 0.85:
                lengths = [int(x) for x in rng.integers(3, 9, 2)]
            else:
                lengths = [int(x) for x in rng.integers(3, 8, 3)]
        elif k == "c4a":
            lengths = [int(x) for x in rng.integers(5, 9, int(rng.integers(2, 5)))]
        elif k == "c4b":
            lengths = [int(x) for x in rng.integers(3, 5, 3)]
        else:
            raise ValueError(kind)
        mods.append((str(i + 1), _strand_seqs(rng, lengths, pair_bias), _strand_struct(lengths)))
    return mods


def write_json(path, modules):
    with open(path, "w") as f:
        json.dump({k: {"sequence": s, "struct2d": ss} for k, s, ss in modules}, f)
    return path


def json_fuzz_modules(n_modules, seed):
    """Deliberately awkward JSON modules: lower case, T, misaligned '&', all bracket families,
    crossing brackets, empty strands, plus invalid entries of every reject class."""
    rng = np.random.default_rng(seed)
    opens, closes = "([{<", ")]}>"
    mods = []
    for i in range(n_modules):
        n_strands = int(rng.integers(1, 5))
        lengths = [int(rng.integers(1, 7)) for _ in range(n_strands)]
        if rng.random() < 0.08 and n_strands > 1:
            # an empty strand, but never the last one: a trailing '&' is dropped by the reference's
            # splitter while struct2d keeps it, and the reference then reads out of bounds
            lengths[int(rng.integers(0, n_strands - 1))] = 0
        total = sum(lengths)
        alphabet = "ACGU" if rng.random() < 0.8 else "ACGUTacgut"
        nts = "".join(alphabet[j] for j in rng.integers(0, len(alphabet), total))
        # random well-formed bracket string over `total` positions, several families, may cross
        ss = ["."] * total
        free = list(range(total))
        rng.shuffle(free)
        n_pairs = int(rng.integers(0, min(4, total // 2) + 1))
        for p in range(n_pairs):
            a, b = sorted((free[2 * p], free[2 * p + 1]))
            fam = int(rng.integers(0, 4))
            ss[a], ss[b] = opens[fam], closes[fam]
        # per-family balance can still underflow when families interleave wrongly: same family
        # pairs placed as (a<b) always balance on their own, so this string is valid.
        seq_parts, ss_parts, at = [], [], 0
        for k in lengths:
            seq_parts.append(nts[at:at + k])
            ss_parts.append("".join(ss[at:at + k]))
            at += k
        seq = "&".join(seq_parts)
        st = "&".join(ss_parts)
        r = rng.random()
        if r < 0.10 and "&" in st:       # misalign the separators of struct2d (same count, other places)
            flat = st.replace("&", "")
            cuts = sorted(rng.choice(np.arange(0, len(flat) + 1), size=st.count("&"), replace=True))
            pieces, prev = [], 0
            for c in cuts:
                pieces.append(flat[prev:c])
                prev = c
            pieces.append(flat[prev:])
            st = "&".join(pieces)
        elif r < 0.14:
            seq = seq + "X"               # 'x' (length mismatch)
        elif r < 0.18 and len(seq) > 0:
            j = int(rng.integers(0, len(seq)))
            if seq[j] != "&":
                seq = seq[:j] + "N" + seq[j + 1:]   # 'r'
        elif r < 0.22 and len(st) > 0:
            j = int(rng.integers(0, len(st)))
            st = st[:j] + ")" + st[j + 1:]      # usually 'n'
        mods.append((f"m{i}", seq, st))
    return mods


# --------------------------------------------------------------------------------------
# CaRNAval RIN files
# --------------------------------------------------------------------------------------
def write_rin_folder(folder, n_modules, seed, invalid_fraction=0.1, max_components=6, min_len=1):
    """Writes <folder>/Subfiles/<i>.txt. 2-max_components components of min_len-6 nt; links
    a<=b over motif nucleotide indices. (Many one-nucleotide components make the number of
    placements, which the reference materialises before filtering, explode with n.)"""
    rng = np.random.default_rng(seed)
    sub = os.path.join(folder, "Subfiles")
    os.makedirs(sub, exist_ok=True)
    for i in range(n_modules):
        n_comp = int(rng.integers(2, max_components + 1))
        lengths = [int(rng.integers(min_len, 7)) for _ in range(n_comp)]
        r = rng.random()
        if r < invalid_fraction / 2:
            lengths = [1, 1, 1]   # too short: 'l'
        while sum(lengths) < 5 and r >= invalid_fraction / 2:
            lengths[int(rng.integers(0, n_comp))] += 1
        total = sum(lengths)
        n_links = int(rng.integers(1, 5))
        if invalid_fraction / 2 <= r < invalid_fraction:
            n_links = 0           # 'x'
        links = ""
        for _ in range(n_links):
            a, b = sorted(int(x) for x in rng.integers(0, total, 2))
            links += f"{a},{b},{'True' if rng.random() < 0.5 else 'False'};"
        lines = ["ntA,ntB,long_range;...", links, "pos;k;seq"]
        at = 0
        for k in lengths:
            seq = "".join(BASES[j] for j in rng.integers(0, 4, k))
            lines.append(f"{at},{at + k - 1};{k};{seq}")
            at += k
        with open(os.path.join(sub, f"{i}.txt"), "w") as f:
            f.write("\n".join(lines))   # no trailing newline, like the installer's last component line + '\n'? see note
            f.write("\n")
    return folder


# --------------------------------------------------------------------------------------
# Rna3Dmotif .desc files
# --------------------------------------------------------------------------------------
def write_desc_folder(folder, n_modules, seed, invalid_fraction=0.1, short_fraction=0.25):
    """Writes <folder>/<name>.desc. Components of 3-8 nt (1-2 nt with probability
    short_fraction), holes of 1-4 unlisted positions inside components, numbering gaps > 5
    between components."""
    rng = np.random.default_rng(seed)
    os.makedirs(folder, exist_ok=True)
    for i in range(n_modules):
        n_comp = int(rng.integers(1, 4))
        pos = int(rng.integers(1, 500))
        bases = []
        for c in range(n_comp):
            k = int(rng.integers(1, 3)) if rng.random() < short_fraction else int(rng.integers(3, 9))
            for j in range(k):
                bases.append((pos, BASES[int(rng.integers(0, 4))]))
                pos += 1 if rng.random() < 0.85 else int(rng.integers(2, 6))
            pos += int(rng.integers(6, 40))
        inter = []
        for (p, nt), (p2, nt2) in zip(bases, bases[1:]):
            if p2 - p == 1:
                inter.append(f"\t({p}_{nt}) C/C ({p2}_{nt2})")
        if len(bases) >= 2:
            (p, nt), (p2, nt2) = bases[0], bases[-1]
            if p2 - p >= 4:
                inter.append(f"\t({p}_{nt}) +/+ ({p2}_{nt2})")
        r = rng.random()
        if r < invalid_fraction * 0.25:
            j = int(rng.integers(0, len(bases)))
            bases[j] = (bases[j][0], "N")                      # unknown nucleotide
        elif r < invalid_fraction * 0.5 and len(bases) >= 3:
            inter.append(f"\t({bases[0][0]}_{bases[0][1]}) C/C ({bases[2][0]}_{bases[2][1]})")   # 'b'
        elif r < invalid_fraction * 0.75 and len(bases) >= 2:
            inter.append(f"\t({bases[0][0]}_{bases[0][1]}) -/- ({bases[1][0]}_{bases[1][1]})")   # 'l'
        elif r < invalid_fraction:
            bases[0] = (0, bases[0][1])                        # '-'
        body = "Bases: " + "".join(f"{p}_{nt}  " for p, nt in bases)
        with open(os.path.join(folder, f"mod{i:05d}.desc"), "w") as f:
            f.write(f"id: {i}\n{body}\n" + "".join(line + "\n" for line in inter))
    return folder
