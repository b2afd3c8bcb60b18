"""DESC stress: 5000 .desc modules (wildcards, delay 5, gates) on a 2000-nt query."""
import sys, time, tempfile, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
seq = synth.sequence(n, 4)
mask = synth.mask_canonical(seq, seed=204, keep=0.3, span=100)
with tempfile.TemporaryDirectory() as tmp:
    path = synth.write_desc_folder(os.path.join(tmp, "desc"), m, 11)
    host = b.HostLibrary.load("descfolder", path)
    t = host.table()
    dev = host.to_device(0)
    q = dev.query([seq], [mask])
    for it in range(3):
        s = dev.search_query(q); st = dev.stats(); ns = s.n_sites; s.close()
    print("descfolder modules", host.n_modules, "units", int(t["n_units"]), "gates", int(t["n_gates"]), "sites", ns, {k: round(st[k], 4) for k in ("count_ms", "scan_ms", "emit_ms")})
