"""ctypes binding of the C ABI declared in include/svfm.h (sview_fmindex_b200/libsvfm.so).

The library is built in-tree by `__graft_entry__.build()` / `make -C sview_fmindex_b200/csrc`.  There is no
CPU fallback: importing this module without the built library, or calling into it without a CUDA device,
fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVFM_LIB_PATH") or os.path.join(_HERE, "libsvfm.so")  # override: developer builds only
BENCH_LIB_PATH = os.path.join(_HERE, "libsvfm_bench.so")  # include/svfm_bench.h: measurement helpers, not the product

SVFM_OK = 0
SVFM_ERR_INVALID_FORMAT = 1
SVFM_ERR_BLOB_SIZE = 2
SVFM_ERR_SYMBOL_COUNT_OVER = 10
SVFM_ERR_TEXT_LENGTH = 11
SVFM_ERR_INVALID_BLOB_SIZE = 12
SVFM_ERR_NOT_ALIGNED = 13
SVFM_ERR_INVALID_CONFIG = 14
SVFM_ERR_BAD_TYPE = 20
SVFM_ERR_EMPTY_PATTERN = 21
SVFM_ERR_TOO_LARGE = 22
SVFM_ERR_NOMEM = 23
SVFM_ERR_BAD_SYMBOL = 24
SVFM_ERR_CAPACITY = 25
SVFM_ERR_BAD_ARG = 26
SVFM_ERR_CUDA = 30

SVFM_TUNE_SORT_MIN = 0
SVFM_TUNE_CHUNK = 1
SVFM_TUNE_SWEEP_MIN = 2
SVFM_TUNE_EXT_BITS = 3
SVFM_TUNE_WORKERS = 4
SVFM_TUNE_ILV = 5
SVFM_TUNE_BUCKET_SORTBACK = 6
SVFM_TUNE_SMALL_MAX = 7
SVFM_TUNE_TEXT = 8
SVFM_TUNE_L2_PERSIST = 9
SVFM_TUNE_OWN_RADIX = 10
SVFM_TUNE_FULL_SA = 11
SVFM_TUNE_SWEEP_OCC = 12
SVFM_TUNE_AUTO = 0xFFFFFFFFFFFFFFFE
SVFM_REVERSED = 1
SVFM_SORTED = 2
SVFM_OFFS32 = 4

ERROR_NAMES = {v: k for k, v in list(globals().items()) if k.startswith("SVFM_ERR_")}


class SvfmType(C.Structure):
    _fields_ = [("pos_bits", C.c_uint32), ("planes", C.c_uint32), ("vec_bits", C.c_uint32), ("encoder", C.c_uint32)]


class SvfmInfo(C.Structure):
    _fields_ = [("type", SvfmType), ("device", C.c_int32), ("symbol_count", C.c_uint32), ("kmer_size", C.c_uint32),
                ("sampling_ratio", C.c_uint32), ("text_len", C.c_uint64), ("suffix_array_len", C.c_uint64),
                ("blocks_len", C.c_uint64), ("sentinel_index", C.c_uint64), ("blob_len", C.c_uint64),
                ("header_size", C.c_uint64), ("off_suffix_array", C.c_uint64), ("off_rank_checkpoints", C.c_uint64),
                ("off_blocks", C.c_uint64)]


# every symbol include/svfm.h declares: (name, restype, argtypes)
_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p
EXPORTS = [
    ("svfm_load", C.c_int, [_vp, C.c_size_t, SvfmType, C.c_int, C.POINTER(_vp), _u64p]),
    ("svfm_load_device", C.c_int, [_vp, C.c_size_t, SvfmType, C.c_int, C.POINTER(_vp), _u64p]),
    ("svfm_free", None, [_vp]),
    ("svfm_index_info", C.c_int, [_vp, C.POINTER(SvfmInfo)]),
    ("svfm_index_memory", C.c_int, [_vp, _u64p]),
    ("svfm_check_blob", C.c_int, [_vp, C.c_size_t, SvfmType, C.POINTER(SvfmInfo), _u64p]),
    ("svfm_count_batch", C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp]),
    ("svfm_locate_batch", C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint64, _u64p]),
    ("svfm_locate_batch_alloc", C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp, C.POINTER(_vp), _u64p]),
    ("svfm_free_positions", None, [_vp]),
    ("svfm_count_batch_packed", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    ("svfm_locate_batch_packed", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint64, _u64p]),
    ("svfm_pack_patterns", C.c_int, [_vp, C.c_uint64, C.c_uint32, _vp, C.c_uint32, _vp]),
    ("svfm_count", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, _u64p]),
    ("svfm_locate", C.c_int, [_vp, _vp, C.c_uint64, C.c_uint32, _vp, C.c_uint64, _u64p]),
    ("svfm_session_create", C.c_int, [_vp, C.POINTER(_vp)]),
    ("svfm_session_destroy", None, [_vp]),
    ("svfm_session_sync", C.c_int, [_vp]),
    ("svfm_session_stream", _vp, [_vp]),
    ("svfm_count_batch_device", C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp]),
    ("svfm_locate_batch_device", C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp, C.POINTER(_vp), _u64p]),
    ("svfm_session_set_timing", C.c_int, [_vp, C.c_int]),
    ("svfm_session_get_timing", C.c_int, [_vp, C.POINTER(C.c_double), _u64p, C.c_int]),
    ("svfm_host_alloc", _vp, [C.c_size_t]),
    ("svfm_host_free", None, [_vp]),
    ("svfm_set_tuning", C.c_int, [C.c_int, C.c_uint64]),
    ("svfm_last_error", C.c_char_p, []),
    ("svfm_launch_count", C.c_uint64, []),
    ("svfm_version", C.c_char_p, []),
    ("svfm_blob_size", C.c_int, [SvfmType, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u64p]),
    ("svfm_build", C.c_int, [SvfmType, _vp, C.c_uint64, C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, C.c_uint64, _u64p]),
    ("svfm_build_device", C.c_int, [SvfmType, _vp, C.c_uint64, C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.c_int, _vp, C.c_uint64, _u64p]),
]

# include/svfm_bench.h (measurement / self-check helpers)
_dp = C.POINTER(C.c_double)
BENCH_EXPORTS = [
    ("svfm_bench_synth_text", C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, C.c_uint32, C.c_uint32, C.c_uint8, _vp]),
    ("svfm_bench_synth_patterns", C.c_int, [_vp, C.c_uint64, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint64, _vp]),
    ("svfm_bench_gather32", C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, _dp, _dp, _vp]),
    ("svfm_bench_verify_locate", C.c_int, [_vp, C.c_uint64, _vp, C.c_uint32, C.c_uint64, _vp, _vp, _vp, C.c_uint32,
                                           _vp, _u64p, _u64p, _vp]),
    ("svfm_bench_count_digest", C.c_int, [_vp, C.c_uint32, C.c_uint64, _u64p, _u64p, _vp]),
    ("svfm_bench_flush_l2", C.c_int, [_vp, C.c_uint64, _vp]),
    ("svfm_bench_last_error", C.c_char_p, []),
]

_lib = None
_bench = None


class _Libs:
    """The product library, plus -- loaded on first use of a svfm_bench_* name -- the measurement helpers."""

    def __init__(self, product):
        self._product = product

    def __getattr__(self, name):
        if name.startswith("svfm_bench_"):
            return getattr(bench_lib(), name)
        return getattr(self._product, name)


def bench_lib() -> C.CDLL:
    global _bench
    if _bench is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise ImportError(f"{BENCH_LIB_PATH} is missing: build it with `make -C sview_fmindex_b200/csrc`")
        B = C.CDLL(BENCH_LIB_PATH)
        for name, restype, argtypes in BENCH_EXPORTS:
            fn = getattr(B, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _bench = B
    return _bench


def lib() -> _Libs:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C sview_fmindex_b200/csrc` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, restype, argtypes in EXPORTS:
            fn = getattr(L, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = _Libs(L)
    return _lib
