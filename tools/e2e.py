"""End-to-end latency of the host-buffer entry point: python tools/e2e.py <workload>..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import biorseo_b200 as b
import bench
for wname in sys.argv[1:]:
    w = bench.make_workload(wname, 0)
    host = b.HostLibrary.from_json_modules(w["modules"])
    dev = host.to_device(0)
    batch = b.HostBatch(w["seqs"], w["masks"], pinned=True)
    for _ in range(3):
        dev.search_host(batch).close()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); s = dev.search_host(batch); t1 = time.perf_counter(); s.close()
        ts.append(t1 - t0)
    print(wname, "search_host ms: median", round(1e3 * float(np.median(ts)), 4), "min", round(1e3 * min(ts), 4), dev.stats()["n_sites"])
