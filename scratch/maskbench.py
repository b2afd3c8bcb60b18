import sys; sys.path.insert(0,'/root/repo')
import torch, numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
host = b.HostLibrary.from_json_modules(synth.json_modules(8, 1, "c1")); dev = host.to_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for n in (2048, 8192, 16000):
    pij = torch.rand((n, n), dtype=torch.float32, device="cuda") * 0.004
    out = torch.empty((n, (n + 31) // 32), dtype=torch.int32, device="cuda")
    for layout, (rs, cs) in (("row", (n, 1)), ("col", (1, n))):
        for _ in range(3): dev.pair_mask_device(pij.data_ptr(), n, rs, cs, 1e-3, out.data_ptr(), stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): dev.pair_mask_device(pij.data_ptr(), n, rs, cs, 1e-3, out.data_ptr(), stream.cuda_stream)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = (n * (n - 1) // 2) * 4 + n * ((n + 31) // 32) * 4
        print(n, layout, round(ms * 1e3, 1), "us", round(nbytes / ms / 1e6, 0), "GB/s", round(nbytes / ms / 1e6 / 6541.5, 3))
