"""End-to-end time of bss_search_sharded (ONE process, the library cut over N GPUs; include/biorseo_b200.h) for N = 1, 2, 4, 8 as far
as the box has GPUs: host buffers in, canonical record stream in pinned host memory out, wall clock around the call. The stream
of every N must equal the N = 1 stream. One JSON line per (workload, N).   python tools/sharded_e2e.py [workload ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import biorseo_b200 as b  # noqa: E402


def run(name, steps=20):
    w = bench.make_workload(name, 0)
    host = b.HostLibrary.from_json_modules(w["modules"])
    batch = b.HostBatch(w["seqs"], w["masks"], pinned=True)
    tested = sum(len(s) for s in w["seqs"]) * int(host.table()["n_units"])
    reference = None
    for n in (1, 2, 4, 8):
        if n > torch.cuda.device_count():
            break
        sh = host.to_devices(n, typical_len=max(len(s) for s in w["seqs"]))
        for _ in range(3):
            sh.search_host(batch).close()
        t0 = time.perf_counter()
        for _ in range(steps):
            s = sh.search_host(batch)
            s.close()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        s = sh.search_host(batch)
        rec = s.records()
        if reference is None:
            reference = rec
        per_gpu = [sh.shard_stats(r)["total_ms"] for r in range(n)]
        print(json.dumps({"workload": name, "describe": w["describe"], "n_gpus": n, "path": "bss_search_sharded (one process, host buffers in, pinned host records out)",
                          "ms_per_search": ms, "placements_per_s": tested / (ms * 1e-3), "sites": s.n_sites, "words": s.n_words,
                          "equal_to_one_gpu": bool(np.array_equal(rec, reference)), "ranges": sh.ranges(),
                          "device_ms_per_gpu": per_gpu}), flush=True)
        s.close()
        sh.close()


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c4", "c4_400k"]):
        run(name)
