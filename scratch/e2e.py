import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import biorseo_b200 as b
import bench
for wname in sys.argv[1:]:
    w = bench.make_workload(wname, 0)
    host = b.HostLibrary.from_json_modules(w["modules"])
    dev = host.to_device(0)
    for _ in range(3):
        dev.search_batch(w["seqs"], w["masks"]).close()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); s = dev.search_batch(w["seqs"], w["masks"]); t1 = time.perf_counter(); s.close(); t2 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1))
    print(wname, "search_batch ms", [round(1e3 * a, 3) for a, _ in ts], "close ms", [round(1e3 * c, 3) for _, c in ts][:3], dev.stats())
