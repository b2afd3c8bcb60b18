"""Reproduces tests/test_gpu_parity.py::test_matches_interpreter_and_reference[seed-n-source] outside pytest (debug helper)."""
import ctypes as C, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from biorseo_b200 import capi
if os.environ.get("BSS_LIB"):
    capi.LIB_PATH = os.path.join(ROOT, "tools", "_build", "lib" + os.environ["BSS_LIB"] + ".so")
import numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
import flat_interp
seed, n, source = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
seq = synth.sequence(n, 100 + seed)
mask = synth.mask_all(n) if seed % 2 else synth.mask_canonical(seq, seed=seed, keep=0.7)
with tempfile.TemporaryDirectory() as tmp:
    path = synth.write_json(os.path.join(tmp, "lib.json"), synth.json_modules(150, seed, "c1"))
    host = b.HostLibrary.load(source, path)
table = host.table()
want = flat_interp.search_table(table, seq, mask)
nU = int(table["n_units"])
units = range(nU) if os.environ.get("EACH") else [None]
l = capi.load()
bad = []
for u in units:
    dev = host.to_device()
    for kv in sys.argv[4:]:
        k, v = kv.split("=")
        if k == "skip":
            l.bss_debug_set_skip(int(v))
        else:
            dev.set_tuning(k, int(v))
    if u is not None:
        l.bss_debug_set_skip(((u + 1) << 8) | 8)     # this unit only, walk but drop the sites
    try:
        if u is None and os.environ.get("SYNC_EACH"):
            import torch; torch.cuda.synchronize()
        s = dev.search(seq, mask)
        rec = s.records()
        if u is None:
            print("equal", np.array_equal(rec, want), rec.size, want.size, dev.stats()["window_items"])
        s.close()
    except Exception as exc:
        print("unit", u, "FAILED", str(exc)[:100])
        bad.append(u)
        break
    dev.close()
if os.environ.get("BSS_LIB", "").startswith("assert"):
    buf = (C.c_uint32 * 32)()
    l.bss_debug_assert_counts.argtypes = [C.POINTER(C.c_uint32)]
    print("assert rc", l.bss_debug_assert_counts(buf), "violations", {i: v for i, v in enumerate(buf) if v})
