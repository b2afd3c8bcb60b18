import sys, time, tempfile, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
seq = synth.sequence(n, 4)
mask = synth.mask_canonical(seq, seed=204, keep=0.3, span=100)
with tempfile.TemporaryDirectory() as tmp:
    for source, path in (("rinfolder", synth.write_rin_folder(os.path.join(tmp, "rin"), m, 12, invalid_fraction=0.0, max_components=6, min_len=1)),):
        host = b.HostLibrary.load(source, path)
        t = host.table()
        dev = host.to_device(0)
        q = dev.query([seq], [mask])
        for it in range(3):
            t0 = time.perf_counter(); s = dev.search_query(q); dt = time.perf_counter() - t0
            st = dev.stats(); ns = s.n_sites; s.close()
        print(source, "modules", host.n_modules, "units", int(t["n_units"]), "gates", int(t["n_gates"]), "sites", ns, "wall ms", round(1e3 * dt, 3),
              {k: round(st[k], 4) for k in ("count_ms", "scan_ms", "emit_ms")})
