"""ctypes bindings of libbiorseo_b200.so (include/biorseo_b200.h, include/biorseo_b200_host.h).

Loading fails loudly when the library has not been built: there is no Python or CPU
implementation of the search to fall back to.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libbiorseo_b200.so")

BSS_OK = 0
BSS_ERR_INVALID, BSS_ERR_NO_DEVICE, BSS_ERR_CUDA, BSS_ERR_LIMIT = -1, -2, -3, -4
BSS_ERR_BAD_SEQUENCE, BSS_ERR_OVERFLOW, BSS_ERR_COUNT, BSS_ERR_NOMEM, BSS_ERR_INTERNAL = -5, -6, -7, -8, -9
KIND_JSON, KIND_RIN, KIND_DESC = 0, 1, 2
C3_COMPONENT_LESS_LINKS, C3_INTERIOR = 0, 1
IDX_SECTIONS = ("var_begin", "var_first", "var_second", "var_c3_terms", "overlap_begin", "overlap_vars", "pair_begin", "pair_partner",
                "row_first_pair")
TUNE = {"group": 0, "units_per_chunk": 1, "heavy_ranks": 2, "list_cap": 3, "window": 4, "window_stage": 5, "window_draw": 6, "pattern_table": 7, "phase_events": 8}
KIND_OF_SOURCE = {"jsonfolder": KIND_JSON, "rinfolder": KIND_RIN, "descfolder": KIND_DESC}

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class BssTable(C.Structure):
    _fields_ = [("n_modules", C.c_uint32), ("n_units", C.c_uint32), ("n_gates", C.c_uint32),
                ("unit_module", u32p), ("unit_delay", u32p), ("unit_gate", u32p),
                ("unit_comp_begin", u32p), ("comp_pat_begin", u32p), ("pattern", C.POINTER(C.c_uint8)),
                ("unit_check_begin", u32p), ("checks", C.POINTER(C.c_int16))]


class BssStats(C.Structure):
    _fields_ = [("placements_tested", C.c_uint64), ("n_sites", C.c_uint64), ("n_words", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint32),
                ("count_ms", C.c_float), ("scan_ms", C.c_float), ("emit_ms", C.c_float), ("total_ms", C.c_float),
                ("window_items", C.c_uint64), ("ordinary_items", C.c_uint64), ("heavy_items", C.c_uint64),
                ("placements_enumerated", C.c_uint64)]


# name -> (restype, argtypes); every symbol declared in include/*.h
SIGNATURES = {
    "bss_abi_version": (C.c_int, []),
    "bss_device_count": (C.c_int, []),
    "bss_last_error": (C.c_char_p, []),
    "bss_library_create": (C.c_int, [C.POINTER(BssTable), C.c_int, C.POINTER(C.c_void_p)]),
    "bss_library_destroy": (None, [C.c_void_p]),
    "bss_library_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bss_library_set_module_base": (C.c_int, [C.c_void_p, C.c_uint32]),
    "bss_library_set_tuning": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "bss_library_units": (C.c_uint32, [C.c_void_p]),
    "bss_library_window_units": (C.c_uint32, [C.c_void_p]),
    "bss_placements_enumerated": (C.c_int, [C.c_void_p, C.c_void_p, u64p]),
    "bss_library_modules": (C.c_uint32, [C.c_void_p]),
    "bss_library_module_components": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "bss_query_create": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, u64p, C.c_void_p, u64p, C.POINTER(C.c_void_p)]),
    "bss_query_destroy": (None, [C.c_void_p]),
    "bss_pair_mask_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_float, C.c_void_p, C.c_void_p]),
    "bss_pair_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_float, C.c_void_p]),
    "bss_query_create_pij": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_float, C.POINTER(C.c_void_p)]),
    "bss_search_query": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bss_search_query_into": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, u64p, u64p]),
    "bss_search_query_into_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
    "bss_search_wait": (C.c_int, [C.c_void_p, u64p, u64p]),
    "bss_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bss_search_batch": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, u64p, C.c_void_p, u64p, C.POINTER(C.c_void_p)]),
    "bss_search_batch_resident": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, u64p, C.c_void_p, u64p, C.POINTER(C.c_void_p)]),
    "bss_sharded_create": (C.c_int, [C.POINTER(BssTable), C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "bss_sharded_destroy": (None, [C.c_void_p]),
    "bss_sharded_gpus": (C.c_uint32, [C.c_void_p]),
    "bss_sharded_range": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_int), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "bss_sharded_library": (C.c_void_p, [C.c_void_p, C.c_uint32]),
    "bss_search_sharded": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, u64p, C.c_void_p, u64p, C.POINTER(C.c_void_p)]),
    "bss_shard_ranges": (C.c_int, [C.POINTER(BssTable), C.c_uint32, C.c_uint32, u32p]),
    "bss_gather_cut": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bss_exchange_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p)]),
    "bss_exchange_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bss_exchange_open": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bss_exchange_open_local": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32]),
    "bss_exchange_block": (C.c_void_p, [C.c_void_p]),
    "bss_exchange_capacity": (C.c_uint64, [C.c_void_p]),
    "bss_exchange_run_async": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "bss_exchange_destroy": (None, [C.c_void_p]),
    "bss_sites_count": (C.c_uint64, [C.c_void_p]),
    "bss_sites_words": (C.c_uint64, [C.c_void_p]),
    "bss_sites_records": (u32p, [C.c_void_p]),
    "bss_sites_device_records": (C.c_void_p, [C.c_void_p]),
    "bss_sites_seq_word_begin": (u64p, [C.c_void_p]),
    "bss_sites_copy_out": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "bss_sites_free": (None, [C.c_void_p]),
    "bss_site_index_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "bss_site_index_count": (C.c_uint64, [C.c_void_p, C.c_int]),
    "bss_site_index_host": (u32p, [C.c_void_p, C.c_int]),
    "bss_site_index_device": (C.c_void_p, [C.c_void_p, C.c_int]),
    "bss_site_index_ms": (C.c_float, [C.c_void_p]),
    "bss_site_index_launches": (C.c_uint32, [C.c_void_p]),
    "bss_site_index_free": (None, [C.c_void_p]),
    "bss_measure_int_issue": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "bss_last_stats": (C.c_int, [C.c_void_p, C.POINTER(BssStats)]),
    "bsh_last_error": (C.c_char_p, []),
    "bsh_library_load": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "bsh_library_new": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bsh_library_save_compiled": (C.c_int, [C.c_void_p, C.c_char_p]),
    "bsh_library_load_compiled": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "bsh_library_mapped_bytes": (C.c_uint64, [C.c_void_p]),
    "bsh_library_kind": (C.c_int, [C.c_void_p]),
    "bsh_library_add_json": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]),
    "bsh_library_seal": (C.c_int, [C.c_void_p]),
    "bsh_library_free": (None, [C.c_void_p]),
    "bsh_library_table": (C.c_int, [C.c_void_p, C.POINTER(BssTable)]),
    "bsh_library_delay": (C.c_uint32, [C.c_void_p]),
    "bsh_library_modules": (C.c_uint32, [C.c_void_p]),
    "bsh_library_seen": (C.c_uint32, [C.c_void_p]),
    "bsh_library_rejected": (C.c_uint32, [C.c_void_p]),
    "bsh_library_reject": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_char)]),
    "bsh_library_module_id": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "bsh_library_module_components": (C.c_uint32, [C.c_void_p, C.c_uint32]),
    "bsh_device_library": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bsh_records_text": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p), u64p]),
    "bsh_text_free": (None, [C.c_void_p]),
    "bsh_fasta_text": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p), u64p]),
}

_lib = None


def load():
    """Returns the loaded shared library, binding every declared symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m biorseo_b200.build` "
                "(or __graft_entry__.build()). There is no fallback implementation.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
