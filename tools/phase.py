"""Per-phase cycle breakdown (debug build with -DBSS_PHASE_TIMING into scratch/libphase.so)."""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from biorseo_b200 import build, capi
if sys.argv[1] == "build":
    srcs = [os.path.join(build.HERE, s) for s in build.CUDA_SOURCES + build.HOST_SOURCES]
    cmd = [build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-DBSS_PHASE_TIMING",
           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(build.HERE, "host"), "-I", os.path.join(build.HERE, "csrc"),
           "-o", os.path.join(ROOT, "tools", "_build", "libphase.so")] + srcs
    os.makedirs(os.path.join(ROOT, "tools", "_build"), exist_ok=True)
    subprocess.run(cmd, check=True); sys.exit(0)
capi.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libphase.so")
import numpy as np
import biorseo_b200 as b
from biorseo_b200 import synth
import bench
names = ["fetch+blob", "gate+match", "thin", "lists", "DP", "walk(all)", "dfs", "tail/other", "tables", "walkA", "walkB"]
for wname in sys.argv[1:]:
    import tempfile
    if wname == "desc":   # 5000 .desc modules on the c4 query
        from biorseo_b200 import synth
        w = bench.make_workload("c4a", 0)
        with tempfile.TemporaryDirectory() as tmp:
            host = b.HostLibrary.load("descfolder", synth.write_desc_folder(os.path.join(tmp, "desc"), 5000, 11))
    else:
        w = bench.make_workload(wname, 0)
        with tempfile.TemporaryDirectory() as tmp:
            host, _ = bench.host_library(w, tmp)
    dev = host.to_device(0)
    q = dev.query(w["seqs"], w["masks"])
    lib = capi.load()
    lib.bss_debug_phase_cycles.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    buf = (C.c_uint64 * 16)()
    for it in range(3):
        lib.bss_debug_phase_cycles(buf, 1)
        s = dev.search_query(q, fetch=False); st = dev.stats(); s.close()
    lib.bss_debug_phase_cycles(buf, 0)
    tot = sum(buf[i] for i in range(8))
    print(wname, {k: round(st[k], 4) for k in ("count_ms", "emit_ms")}, "total Mcycles", round(tot / 1e6, 1))
    for i, nm in enumerate(names):
        print(f"   {nm:12s} {buf[i]/1e6:10.1f} Mcyc  {100*buf[i]/max(tot,1):5.1f}%")
