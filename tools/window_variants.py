"""Device time of one search per workload and kernel variant (ordinary passes / window pass / window pass with
the query tables staged in shared memory). Prints one JSON line per (workload, variant): the library's own
CUDA-event times of the count / scan / emit sides, the step time between two events on the stream with L2
flushed in between, and the routing counters.   python tools/window_variants.py [workload ...]"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import biorseo_b200 as b  # noqa: E402

VARIANTS = [("ordinary", {"window": 0}), ("window", {}), ("window_no_pattern_table", {"pattern_table": 0}), ("window_staged", {"window_stage": 1}), ("window_no_phase_events", {"phase_events": 0})]
if os.environ.get("DRAW_SWEEP"):
    VARIANTS = [(f"draw{d}", {"window_draw": d}) for d in (1, 2, 4, 8, 16, 32)]


def run(name, steps=20):
    w = bench.make_workload(name, 0)
    with tempfile.TemporaryDirectory() as tmp:
        host, _ = bench.host_library(w, tmp)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    reference = None
    for label, tuning in VARIANTS:
        dev = host.to_device(0)
        for knob, value in tuning.items():
            dev.set_tuning(knob, value)
        dev.set_stream(stream.cuda_stream)
        query = dev.query(w["seqs"], w["masks"])
        rc, n_words, n_sites = dev.search_query_into(query, 0, 0)
        records = torch.empty(max(n_words, 1), dtype=torch.int32, device="cuda")
        ms, kernel = [], {"count_ms": [], "scan_ms": [], "emit_ms": []}
        for i in range(steps + 3):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            dev.search_query_into_async(query, records.data_ptr(), records.numel())
            e1.record(stream)
            rc, nw, ns = dev.search_wait()
            assert rc == 0 and nw == n_words
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(e0.elapsed_time(e1))
                st = dev.stats()
                for k in kernel:
                    kernel[k].append(st[k])
        got = records[:n_words].cpu().numpy()
        if reference is None:
            reference = got
        st = dev.stats()
        print(json.dumps({"workload": name, "variant": label, "step_ms": float(np.median(ms)), **{k: float(np.median(v)) for k, v in kernel.items()},
                          "sites": n_sites, "words": n_words, "equal_to_first_variant": bool(np.array_equal(got, reference)),
                          "window_units": dev.n_window_units, "units": dev.n_units, "window_items": st["window_items"],
                          "ordinary_items": st["ordinary_items"], "heavy_items": st["heavy_items"], "launches": st["kernel_launches"]}), flush=True)
        query.close()
        dev.close()


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c4", "c4b", "c4a", "c4_desc", "c1", "c2"]):
        run(name)
