"""Insertion-variable index build on an output-heavy case: python tools/index_phase.py [units] [repeat]
(run it under `ncu --metrics gpu__time_duration.sum` for the per-kernel times)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
units = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 5
seq = synth.sequence(2000, 4)
host = b.HostLibrary.from_json_modules(synth.json_modules(units, 104, "c4b"))
dev = host.to_device(0)
q = dev.query([seq], [synth.mask_all(2000)])
sites = dev.search_query(q, fetch=False)
for i in range(repeat):
    t0 = time.perf_counter()
    idx = dev.site_index(q, sites.device_pointer(), 0, fetch=False)
    t1 = time.perf_counter()
    print(f"sites {sites.n_sites} vars {idx.count('var_first')} entries {idx.count('overlap_vars')}: device {idx.build_ms:.3f} ms, host call {1e3 * (t1 - t0):.3f} ms, "
          f"{idx.kernel_launches} launches")
    idx.close()
