"""biorseo_b200 — B200-native motif insertion-site search for Biorseo.

The product is the shared library (CUDA kernels for sm_100a behind a C ABI, C++ host glue
with the reference's Motif / Pool interfaces). This Python package is only the binding
used by the test-suite and bench.py.
"""
from .api import (DeviceLibrary, HostBatch, HostLibrary, Query, SearchError, ShardedLibrary, Sites, canonical_pair_mask, pack_mask, read_fasta)  # noqa: F401
