"""Test helper: plain-Python restatement of what the MOIP constructor derives from insertion_sites_
(the insertion-variable index of include/biorseo_b200.h, bss_site_index_build), and the constraints
that follow from it, in a normal form that can be compared with the reference's own constraint
builder (oracle/_ref `model` mode: cppsrc/MOIP.cpp:302-328 and :507-694 run unmodified against a
recording solver shim). Small inputs only. Never used by the product.

A site is the oracle's dump line 'identifier\\tfirst-second ...\\ta-b ...' (components, links).
"""
import numpy as np


def parse_site(line):
    ident, comps, links = (line.split("\t") + ["", ""])[:3]
    comps = [tuple(int(x) for x in c.split("-")) for c in comps.split(" ") if c]
    links = [tuple(int(x) for x in l.split("-")) for l in links.split(" ") if l]
    return ident, comps, links


def allowed(mask, n, u, v):
    """MOIP::allowed_basepair (cppsrc/MOIP.cpp:896-910) on the packed mask: bit (min, max)."""
    a, b = min(u, v), max(u, v)
    if a < 0 or b >= n or a == b:
        return False
    return bool((int(mask[a, b >> 5]) >> (b & 31)) & 1)


def build(site_lines, n, mask, c3_rule):
    """The index sections, by the definitions (cppsrc/MOIP.cpp:74-90, 302-328, 507-586).
    c3_rule 0: jsonfolder / rinfolder, 1: descfolder."""
    mask = np.asarray(mask, dtype=np.uint32).reshape(n, -1) if n else np.zeros((0, 1), np.uint32)
    partners = [[v for v in range(n) if allowed(mask, n, u, v)] for u in range(n)]
    pair_begin = np.concatenate([[0], np.cumsum([len(p) for p in partners])]).astype(np.uint32)
    row_pairs = [sum(1 for v in partners[u] if v > u) for u in range(n)]
    row_first_pair = np.concatenate([[0], np.cumsum(row_pairs)]).astype(np.uint32)
    var_begin, var_first, var_second, var_c3 = [0], [], [], []
    overlap = [[] for _ in range(n)]
    for line in site_lines:
        _, comps, links = parse_site(line)
        linkset = {frozenset(l) for l in links}
        for first, second in comps:
            v = len(var_first)
            k = second - first + 1
            var_first.append(first)
            var_second.append(second & 0xFFFFFFFF)
            if c3_rule == 0:
                count = sum(1 for u in range(first, second + 1) for w in partners[u] if frozenset((u, w)) not in linkset)
            else:
                count = sum(len(partners[u]) for u in range(first + 1, second))
            var_c3.append(count)
            for u in range(first, first + k):
                overlap[u].append(v)
        var_begin.append(len(var_first))
    u32 = lambda x: np.asarray(x, dtype=np.uint32).reshape(-1)
    return {
        "var_begin": u32(var_begin), "var_first": u32(var_first), "var_second": u32(var_second), "var_c3_terms": u32(var_c3),
        "overlap_begin": np.concatenate([[0], np.cumsum([len(o) for o in overlap])]).astype(np.uint32),
        "overlap_vars": u32([v for o in overlap for v in o]),
        "pair_begin": pair_begin, "pair_partner": u32([v for p in partners for v in p]), "row_first_pair": row_first_pair,
    }


def _num(x):
    return f"{float(x) + 0.0:g}"   # + 0.0 turns -0.0 into 0.0


def _norm(op, rhs, terms):
    return (op, _num(rhs), tuple(sorted((name, _num(c)) for c, name in terms)))


def constraints_from_index(index, site_lines, n, mask, c3_rule):
    """The constraints of define_problem_constraints that involve an insertion variable, built the way a
    MOIP constructor that consumes the index would build them: c3 from the variables table + the pair
    adjacency (+ the site's links), c4 from the overlap lists, c5 / c6 from the sites. Sorted normal forms."""
    mask = np.asarray(mask, dtype=np.uint32).reshape(n, -1)
    vb, vf, vs, vc3 = index["var_begin"], index["var_first"], index["var_second"], index["var_c3_terms"]
    pb, pp = index["pair_begin"], index["pair_partner"]
    assert len(vb) == len(site_lines) + 1
    names, sites = [], []
    for x, line in enumerate(site_lines):
        ident, comps, links = parse_site(line)
        assert int(vb[x + 1]) - int(vb[x]) == len(comps)
        sites.append((comps, links))
        for j in range(len(comps)):
            v = int(vb[x]) + j
            names.append(f"C<{line}>,{j}[{int(vf[v])},{int(np.int32(vs[v]))}]")
    y = lambda u, v: f"y{min(u, v)},{max(u, v)}"
    out = []
    for x, (comps, links) in enumerate(sites):
        linkset = {frozenset(l) for l in links}
        for j, (first, second) in enumerate(comps):
            v = int(vb[x]) + j
            k = second - first + 1
            if c3_rule == 0:
                ys = [y(u, int(w)) for u in range(first, second + 1) for w in pp[int(pb[u]):int(pb[u + 1])] if frozenset((u, int(w))) not in linkset]
                coef = k
            else:
                ys = [y(u, int(w)) for u in range(first + 1, second) for w in pp[int(pb[u]):int(pb[u + 1])]]
                coef = k - 2
            assert len(ys) == int(vc3[v]), f"c3 term count of variable {v}: index says {int(vc3[v])}, lists give {len(ys)}"
            if ys:
                out.append(_norm("<", coef, [(coef, names[v])] + [(1, t) for t in ys]))
    ob, ov = index["overlap_begin"], index["overlap_vars"]
    for u in range(n):
        covering = [int(v) for v in ov[int(ob[u]):int(ob[u + 1])]]
        assert covering == sorted(covering)
        if len(covering) > 1:
            out.append(_norm("<", 1, [(1, names[v]) for v in covering]))
    for x, (comps, links) in enumerate(sites):
        v0 = int(vb[x])
        if len(comps) > 1:
            out.append(_norm("=", 0, [(1, names[v0 + j]) for j in range(1, len(comps))] + [(-(len(comps) - 1), names[v0])]))
        if c3_rule == 0:
            for j, (first, second) in enumerate(comps):
                ax, terms = 0, []
                for a, b in links:
                    if first <= a <= second:
                        ax += 1
                        if not allowed(mask, n, a, b):
                            break
                        terms.append((1, y(a, b)))
                if ax:
                    out.append(_norm(">", 0, terms + [(-ax, names[v0 + j])]))
        else:
            terms = [(1, names[v0])]
            if allowed(mask, n, comps[0][0], comps[-1][1]):
                terms.append((-1, y(comps[0][0], comps[-1][1])))
            out.append(_norm("<", 0, terms))
            for j in range(len(comps) - 1):
                terms = [(1, names[v0 + j])]
                if allowed(mask, n, comps[j][1], comps[j + 1][0]):
                    terms.append((-1, y(comps[j][1], comps[j + 1][0])))
                out.append(_norm("<", 0, terms))
    return sorted(out)


def parse_model_dump(text):
    """oracle/_ref `model` output -> (site lines in the reference's order, sorted constraint normal forms with the
    C_{x,i,p} names rewritten from the (race-dependent) site number x to the site itself)."""
    sites, raw = [], []
    for line in text.split("\n"):
        if line.startswith("S\t"):
            sites.append(line[2:])
        elif line.startswith("K "):
            raw.append(line)
    out = []
    for line in raw:
        _, op, rhs, *terms = line.split(" ")
        parsed = []
        for t in terms:
            coef, name = t.split("*", 1)
            if name.startswith("C"):
                x, rest = name[1:].split(",", 1)
                name = f"C<{sites[int(x)]}>,{rest}"
            parsed.append((float(coef), name))
        out.append(_norm(op, float(rhs), parsed))
    return sites, sorted(out)
