"""ctypes view of libkcfgpu.so (the C ABI declared in include/kcf_b200.h).

There is no fallback of any kind: a missing library or a machine without a B200-class CUDA device
raises.  Build the library with `python __graft_entry__.py` (or `make -C kcftools_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KCF_LIB_PATH") or os.path.join(_HERE, "libkcfgpu.so")  # KCF_LIB_PATH: tuning builds

RESULT_DTYPE = np.dtype([("total_kmers", "<i4"), ("eff_len", "<i4"), ("obs", "<i4"), ("variations", "<i4"),
                         ("inner", "<i4"), ("left", "<i4"), ("right", "<i4"), ("_pad", "<i4"),
                         ("kmer_count_sum", "<i8"), ("score", "<f8")])
WINDOW_DTYPE = np.dtype([("first_seg", "<u4"), ("n_segs", "<u4")])
SEGMENT_DTYPE = np.dtype([("seq_id", "<i4"), ("start0", "<i4"), ("len", "<i4")])
CELL_DTYPE = np.dtype([("obs", "<i4"), ("variations", "<i4"), ("inner", "<i4"), ("left", "<i4"), ("right", "<i4"), ("ibs", "<i4"),
                       ("kmer_count", "<i8"), ("score", "<f8")])  # kcf_cell_t
assert CELL_DTYPE.itemsize == 40
assert RESULT_DTYPE.itemsize == 48 and WINDOW_DTYPE.itemsize == 8 and SEGMENT_DTYPE.itemsize == 12

STATUS = {0: "KCF_OK", -1: "KCF_ERR_CUDA", -2: "KCF_ERR_IO", -3: "KCF_ERR_DB_FORMAT", -4: "KCF_ERR_UNSUPPORTED",
          -5: "KCF_ERR_ARG", -6: "KCF_ERR_RANGE", -7: "KCF_ERR_FASTA", -8: "KCF_ERR_WEIGHTS", -9: "KCF_ERR_NOMEM",
          -10: "KCF_ERR_DB_ORDER"}


class HostSeq(C.Structure):  # kcf_host_seq_t
    _fields_ = [("bytes", C.c_void_p), ("n_bytes", C.c_uint64), ("line_bases", C.c_uint32), ("line_width", C.c_uint32),
                ("seq_len", C.c_uint64)]


class DbInfo(C.Structure):
    _fields_ = [("kmer_length", C.c_int32), ("lut_prefix_length", C.c_int32), ("signature_length", C.c_int32),
                ("counter_size", C.c_int32), ("both_strands", C.c_int32), ("min_count", C.c_int32),
                ("max_count", C.c_int32), ("n_bins", C.c_int32), ("total_kmers", C.c_int64),
                ("resident_kmers", C.c_int64), ("unreachable_kmers", C.c_int64), ("stash_kmers", C.c_int64),
                ("table_bytes", C.c_int64), ("n_buckets", C.c_int64), ("load_seconds", C.c_double),
                ("elsewhere_kmers", C.c_int64), ("load_phase_s", C.c_double * 4)]


# every symbol include/kcf_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "kcf_init": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "kcf_shutdown": (None, [_P]),
    "kcf_last_error": (C.c_char_p, [_P]),
    "kcf_stream": (_P, [_P]),
    "kcf_host_alloc": (C.c_int, [_P, C.c_uint64, C.POINTER(_P)]),
    "kcf_host_free": (None, [_P, _P]),
    "kcf_db_open": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(_P)]),
    "kcf_db_open_mem": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "kcf_db_info": (C.c_int, [_P, C.POINTER(DbInfo)]),
    "kcf_db_close": (None, [_P]),
    "kcf_set_load_factor": (C.c_int, [_P, C.c_double]),
    "kcf_set_minimizer_length": (C.c_int, [_P, C.c_int]),
    "kcf_db_count": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "kcf_ref_add": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_int)]),
    "kcf_ref_add_async": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_int)]),
    "kcf_ref_sync": (C.c_int, [_P]),
    "kcf_ref_clear": (C.c_int, [_P]),
    "kcf_screen": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_uint64, C.c_int32, C.POINTER(C.c_double), _P]),
    "kcf_set_upload_piece": (C.c_int, [_P, C.c_uint64]),
    "kcf_shard_windows": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64, C.c_int, _P]),
    "kcf_screen_sharded": (C.c_int, [_P, _P, C.c_int, _P, C.c_uint32, _P, C.c_uint64, _P, C.c_uint64, C.c_int32, C.POINTER(C.c_double), _P]),
    "kcf_plan_create": (C.c_int, [_P, C.c_int32, _P, C.c_uint64, _P, C.c_uint64, C.POINTER(_P)]),
    "kcf_plan_run": (C.c_int, [_P, _P, _P, C.c_int32, C.POINTER(C.c_double)]),
    "kcf_plan_fetch": (C.c_int, [_P, _P, _P]),
    "kcf_plan_destroy": (None, [_P]),
    "kcf_plan_stats": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "kcf_window_counts": (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "kcf_set_partition": (C.c_int, [_P, C.c_int, C.c_int]),
    "kcf_xchg_extract": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_int, _P, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "kcf_xchg_lookup": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P]),
    "kcf_xchg_fold": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P, C.c_uint64, C.c_int32]),
    "kcf_xg_create": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_uint64, C.POINTER(_P)]),
    "kcf_xg_export": (C.c_int, [_P, _P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "kcf_xg_connect": (C.c_int, [_P, _P, _P]),
    "kcf_xg_destroy": (None, [_P]),
    "kcf_xg_send": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_uint64]),
    "kcf_xg_answer": (C.c_int, [_P, _P, _P]),
    "kcf_xg_fold": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_int32]),
    "kcf_xg_status": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "kcf_xg_pipeline": (C.c_int, [_P, C.c_uint32]),
    "kcf_xg_join": (C.c_int, [_P, _P]),
    "kcf_plan_finalize": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "kcf_cohort_create": (C.c_int, [_P, C.c_uint64, C.c_uint32, _P, _P, C.POINTER(_P)]),
    "kcf_cohort_destroy": (None, [_P]),
    "kcf_cohort_add_plan": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, _P]),
    "kcf_cohort_set_sample": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "kcf_cohort_scores": (C.c_int, [_P, _P, C.POINTER(C.c_double)]),
    "kcf_cohort_find_ibs": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.c_int, C.c_int32, C.c_float]),
    "kcf_cohort_genotypes": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P]),
    "kcf_cohort_fetch": (C.c_int, [_P, _P, C.c_uint32, _P, _P, _P]),
    "kcf_db_line_histogram": (C.c_int, [_P, _P]),
    "kcf_scan_owned": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_int32, _P, _P]),
    "kcf_scan_fold": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, _P]),
    "kcf_measure_random_sector_gbps": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "kcf_measure_random_line_rate": (C.c_int, [_P, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
    "kcf_set_profiling": (C.c_int, [_P, C.c_int]),
    "kcf_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "kcf_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """dlopen libkcfgpu.so and bind every declared symbol; raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                               "(there is no CPU or PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class KcfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.msg = msg
