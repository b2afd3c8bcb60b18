"""Per-module difference between the GPU search of a golden case and its expected answer (debug helper)."""
import collections, ctypes as C, os, sys, tempfile, pathlib
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import capi
if os.environ.get("BSS_LIB") == "assert":
    capi.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libassert.so")
import numpy as np
import biorseo_b200 as b
import cases, flat_interp
name = sys.argv[1]
case = [c for c in cases.CASES if c["name"] == name][0]
with tempfile.TemporaryDirectory() as tmp:
    lib = cases.materialise(case, pathlib.Path(tmp))
    host = b.HostLibrary.load(case["source"], lib)
    table = host.table()
    seq = case["seq"]
    mask = cases.mask_of(case["mask"], seq)
    want = flat_interp.search_table(table, seq, mask)
    comps = host.module_components()
    def per_module(rec):
        out, at = collections.Counter(), 0
        while at < rec.size:
            out[int(rec[at])] += 1
            at += 1 + int(comps[rec[at]])
        return out
    w = per_module(want)
    for reps in range(int(os.environ.get('REPS', '1'))):
        dev = host.to_device()
        for kv in sys.argv[2:]:
            k, v = kv.split("=")
            if k == "skip":
                capi.load().bss_debug_set_skip(int(v))
            elif k == "grid":
                capi.load().bss_debug_set_launch(int(v), int(os.environ.get("SMEM_EXTRA", "0")))
            else:
                dev.set_tuning(k, int(v))
        q = dev.query([seq], [mask])
        try:
            s = dev.search_query(q)
        except Exception as exc:
            print("FAILED", exc)
            l = capi.load()
            buf = (C.c_uint32 * 32)()
            l.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
            print("   assert rc", l.bss_debug_assert_counts(buf), "violations", {i: v for i, v in enumerate(buf) if v})
            break
        rec = s.records()
        g = per_module(rec)
        diff = {m: (g.get(m, 0), w.get(m, 0)) for m in set(g) | set(w) if g.get(m, 0) != w.get(m, 0)}
        print(name, sys.argv[2:], "words", rec.size, want.size, "equal", np.array_equal(rec, want), "modules that differ (gpu, want):", diff, dev.stats()["window_items"])
        def sites_of(rec, m):
            out, at = [], 0
            while at < rec.size:
                wdt = 1 + int(comps[rec[at]])
                if int(rec[at]) == m:
                    out.append(tuple(int(x) for x in rec[at + 1:at + wdt]))
                at += wdt
            return out
        for m in sorted(diff)[:4]:
            gs, ws = sites_of(rec, m), sites_of(want, m)
            print("   module", m, "missing", sorted(set(ws) - set(gs))[:40], "extra", sorted(set(gs) - set(ws))[:10])
            u = [i for i in range(int(table["n_units"])) if int(table["unit_module"][i]) == m]
            for i in u:
                pats, checks, delay = flat_interp._unit(table, i)
                print("   unit", i, "module", m, host.module_id(m), pats, checks, delay)
        if os.environ.get("BSS_LIB") == "assert":
            l = capi.load()
            nU = int(table["n_units"])
            counts = np.zeros(nU, np.uint32); rows = np.zeros((nU, 32), np.uint32); alive = np.zeros(nU, np.uint32)
            l.bss_debug_counts.argtypes = [C.c_void_p] * 2 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
            l.bss_debug_counts(dev._h, counts.ctypes.data_as(C.c_void_p), nU, rows.ctypes.data_as(C.c_void_p), nU, alive.ctypes.data_as(C.c_void_p))
            wc = per_module(want)
            mod = table["unit_module"]
            print("   counts per unit (gpu/want):", [(u, int(counts[u]), wc.get(int(mod[u]), 0)) for u in range(nU) if int(counts[u]) != wc.get(int(mod[u]), 0)])
            st = dev.stats()
            print("   alive_w", alive[:8].tolist(), "rows of the first alive window items:", [(int(alive[i]), rows[i][rows[i] > 0].tolist(), np.nonzero(rows[i])[0].tolist()) for i in range(3)])
            buf = (C.c_uint32 * 32)()
            l.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
            print("   assert rc", l.bss_debug_assert_counts(buf), "violations", {i: v for i, v in enumerate(buf) if v})
        s.close(); q.close(); dev.close()
