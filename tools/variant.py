"""A/B timing of two builds of the library: python tools/variant.py <lib.so> <workload>..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biorseo_b200 import capi
capi.LIB_PATH = sys.argv[1]
import numpy as np
import biorseo_b200 as b
import bench
for wname in sys.argv[2:]:
    w = bench.make_workload(wname, 0)
    host = b.HostLibrary.from_json_modules(w["modules"])
    dev = host.to_device(0)
    q = dev.query(w["seqs"], w["masks"])
    cm, em = [], []
    for it in range(14):
        s = dev.search_query(q, fetch=False); st = dev.stats(); s.close()
        if it >= 4: cm.append(st["count_ms"]); em.append(st["emit_ms"])
    print(os.path.basename(sys.argv[1]), wname, "count_ms", round(float(np.median(cm)), 4), "emit_ms", round(float(np.median(em)), 4))
