"""Builds tools/_build/lib<name>.so with extra nvcc flags (debug helper): python tools/gpu/build_variant.py name -DFLAG ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from biorseo_b200 import build
name, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, "tools", "_build"), exist_ok=True)
srcs = [os.path.join(build.HERE, s) for s in build.CUDA_SOURCES + build.HOST_SOURCES]
subprocess.run([build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared"] + flags +
               ["-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(build.HERE, "host"), "-I", os.path.join(build.HERE, "csrc"),
                "-o", os.path.join(ROOT, "tools", "_build", f"lib{name}.so")] + srcs, check=True)
