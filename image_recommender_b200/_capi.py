"""ctypes binding of include/b2k.h (the C ABI of libb2k.so).

The library is the only compute path of this package: if it is missing the import of
`image_recommender_b200.index` fails loudly; there is no Python/numpy/torch fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libb2k.so"

B2K_MAX_TABLES = 8
B2K_MAX_K = 1024
B2K_LIST = 32

OPT_PATH, OPT_RERANK, OPT_FORCE_EXACT, OPT_SCAN_MAX_B, OPT_SPLITS, OPT_TC_PAIR, OPT_SEED, OPT_TIGHTEN, OPT_COLLECT, OPT_INLINE_SEED = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
OPT_FUSED_TAIL, OPT_TN, OPT_SAMPLE_WAVE = 11, 12, 13
PATH_AUTO, PATH_SCAN, PATH_TC = 0, 1, 2
E_INVALID, E_CAPACITY, E_IO, E_NODEVICE, E_NOMEM, E_UNSUPPORTED, E_PEER = -1, -2, -3, -4, -5, -6, -7


class B2KError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"b2k status {status}: {message}")
        self.status = status


class Stats(C.Structure):
    _fields_ = [("path", C.c_int32), ("n_splits", C.c_int32), ("cand_slots", C.c_int32),
                ("n_uncertified", C.c_int32), ("eps_max", C.c_float), ("err_max", C.c_float),
                ("norm_max", C.c_float), ("launches", C.c_int32), ("score_ms", C.c_float),
                ("tail_ms", C.c_float), ("n_queries", C.c_int32), ("n_candidates", C.c_int32),
                ("n_saturated", C.c_int32), ("n_timed", C.c_int32)]


class Synth(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_clusters", C.c_int32), ("sigma", C.c_float),
                ("abs_mask", C.c_uint32), ("total_rows", C.c_int64)]


# name -> (restype, argtypes); every symbol include/b2k.h declares
SIGNATURES = {
    "b2k_last_error": (C.c_char_p, []),
    "b2k_abi_version": (C.c_int32, []),
    "b2k_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "b2k_create": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.c_int64, C.c_int32, C.c_int64,
                             C.POINTER(C.c_void_p)]),
    "b2k_destroy": (None, [C.c_void_p]),
    "b2k_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "b2k_capacity": (C.c_int64, [C.c_void_p]),
    "b2k_reset": (C.c_int, [C.c_void_p]),
    "b2k_add": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64]),
    "b2k_add_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64, C.c_void_p]),
    "b2k_stage_open": (C.c_int, [C.c_void_p, C.c_int64]),
    "b2k_stage_rows": (C.c_int64, [C.c_void_p]),
    "b2k_stage_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "b2k_stage_wait": (C.c_int, [C.c_void_p, C.c_int32]),
    "b2k_stage_commit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "b2k_stage_close": (C.c_int, [C.c_void_p]),
    "b2k_ingest_sqlite": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64,
                                    C.POINTER(C.c_int64)]),
    "b2k_ingest_sqlite_mt": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                       C.c_int64, C.POINTER(C.c_int64)]),
    "b2k_stage_open_n": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32]),
    "b2k_parse_f32_blob": (C.c_int, [C.c_char_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "b2k_table_dims": (C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    "b2k_ntotal": (C.c_int64, [C.c_void_p]),
    "b2k_dim": (C.c_int32, [C.c_void_p]),
    "b2k_dim_padded": (C.c_int32, [C.c_void_p]),
    "b2k_base_offset": (C.c_int64, [C.c_void_p]),
    "b2k_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                             C.c_void_p]),
    "b2k_search_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_search_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_prep_groups_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                         C.c_void_p]),
    "b2k_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "b2k_set_option": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "b2k_merge_topk_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                        C.c_void_p]),
    "b2k_xchg_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    "b2k_xchg_destroy": (None, [C.c_void_p]),
    "b2k_xchg_handle": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2k_xchg_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b2k_xchg_skip": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b2k_xchg_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "b2k_xchg_merge": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_xchg_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)]),
    "b2k_group_create": (C.c_int, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    "b2k_group_destroy": (None, [C.c_void_p]),
    "b2k_group_size": (C.c_int32, [C.c_void_p]),
    "b2k_group_load": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b2k_group_set_shard": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "b2k_group_shard": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "b2k_group_ntotal": (C.c_int64, [C.c_void_p]),
    "b2k_group_dim": (C.c_int32, [C.c_void_p]),
    "b2k_group_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_group_put_queries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "b2k_group_run": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "b2k_group_get_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_group_search_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "b2k_group_last_run_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "b2k_normalize_l2": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "b2k_save": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "b2k_save_shard": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "b2k_load": (C.c_int, [C.c_char_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "b2k_file_info": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b2k_load_ids": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "b2k_get_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "b2k_fill_synthetic": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Synth)]),
    "b2k_synth_queries_device": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(Synth), C.c_uint64,
                                           C.c_float, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen libb2k.so and bind every entry point; raises if the extension was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -m image_recommender_b200.build_ext, or __graft_entry__.build()). "
            "image_recommender_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load_library().b2k_last_error()
        raise B2KError(status, msg.decode("utf-8", "replace") if msg else "")
