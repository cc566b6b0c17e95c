"""ctypes front end of the CPU ORACLE (oracle/libfm_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (sview_fmindex_b200) never does.

The class and method names follow the reference (baku4/sview-fmindex) public API:
`FmIndexBuilder::new/blob_size/build` (builder/mod.rs:63-264), `FmIndex::load/count/locate`
(load_from_blob.rs:28, locate/with_slice.rs:5-18), `count_rev_iter/locate_rev_iter`
(locate/with_rev_iter.rs:5-18), `EncodingTable::from_symbols[_with_wildcard]`
(components/text_encoder/text_encoders/encoding_table.rs:15-34).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfm_oracle.so")


def build_oracle(force: bool = False) -> str:
    """Compile oracle/libfm_oracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("fm_oracle.c", "fm_query.inc", "fm_oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libfm_oracle.so"])
    return _LIB_PATH


class OraType(C.Structure):
    _fields_ = [("pos_bits", C.c_uint32), ("planes", C.c_uint32), ("vec_bits", C.c_uint32), ("encoder", C.c_uint32)]


class OraLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "align", "off_encoder", "off_count_header", "off_sa_header", "off_bwm_header", "header_size",
        "off_count_array", "off_kmer_multiplier", "off_kmer_count_table", "off_suffix_array",
        "off_sentinel_index", "off_rank_checkpoints", "off_blocks", "total_size")] + [
        ("symbol_count", C.c_uint32), ("kmer_size", C.c_uint32), ("count_array_len", C.c_uint32),
        ("kmer_multiplier_len", C.c_uint32), ("kmer_count_table_len", C.c_uint64),
        ("sampling_ratio", C.c_uint32), ("suffix_array_len", C.c_uint64),
        ("rank_checkpoints_len", C.c_uint64), ("blocks_len", C.c_uint64)]


ORA_OK = 0
ORA_ERR_INVALID_FORMAT = 1
ORA_ERR_BLOB_SIZE = 2
ORA_ERR_SYMBOL_COUNT_OVER = 10
ORA_ERR_INVALID_BLOB_SIZE = 12
ORA_ERR_INVALID_CONFIG = 14
ORA_ERR_EMPTY_PATTERN = 21

_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_LIB_PATH)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.ora_encoding_table.restype = C.c_uint32
        L.ora_encoding_table.argtypes = [C.c_char_p, u32p, C.c_uint32, C.c_int, u8p]
        L.ora_builder_layout.restype = C.c_int
        L.ora_builder_layout.argtypes = [OraType, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.POINTER(OraLayout), u64p]
        L.ora_kmer_size_for_max_memory.restype = C.c_uint32
        L.ora_kmer_size_for_max_memory.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.ora_build.restype = C.c_int
        L.ora_build.argtypes = [OraType, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint32,
                                C.c_uint32, C.c_void_p, C.c_uint64, u64p]
        L.ora_load.restype = C.c_int
        L.ora_load.argtypes = [C.c_void_p, C.c_uint64, OraType, C.POINTER(C.c_void_p), u64p]
        L.ora_free.restype = None
        L.ora_free.argtypes = [C.c_void_p]
        L.ora_index_layout.restype = C.POINTER(OraLayout)
        L.ora_index_layout.argtypes = [C.c_void_p]
        L.ora_text_len.restype = C.c_uint64
        L.ora_text_len.argtypes = [C.c_void_p]
        L.ora_count.restype = C.c_int
        L.ora_count.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, u64p]
        L.ora_locate.restype = C.c_int
        L.ora_locate.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint64, u64p]
        L.ora_pos_range.restype = C.c_int
        L.ora_pos_range.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, u64p, u64p]
        L.ora_count_batch.restype = C.c_int
        L.ora_count_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        L.ora_locate_batch.restype = C.c_int
        L.ora_locate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_uint32, u64p, C.c_int]
        L.ora_suffix_array.restype = C.c_int
        L.ora_suffix_array.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code, detail=(0, 0)):
        super().__init__(f"oracle error {code} detail={tuple(detail)}")
        self.code = code
        self.detail = tuple(detail)


@dataclass(frozen=True)
class IndexType:
    """The (P, B, E) triple of FmIndex<'a, P, B, E>: e.g. IndexType(32, 3, 64, True) = <u32, Block3<u64>, EncodingTable>."""
    pos_bits: int = 32
    planes: int = 3
    vec_bits: int = 64
    encoding_table: bool = True

    def c(self) -> OraType:
        return OraType(self.pos_bits, self.planes, self.vec_bits, 1 if self.encoding_table else 0)

    @property
    def pos_dtype(self):
        return np.uint32 if self.pos_bits == 32 else np.uint64


def encoding_table(symbols, with_wildcard: bool = False):
    """EncodingTable::from_symbols / from_symbols_with_wildcard -> (table[256] uint8, symbol_count)."""
    groups = [bytes(s) for s in symbols]
    offs = np.zeros(len(groups) + 1, dtype=np.uint32)
    offs[1:] = np.cumsum([len(g) for g in groups])
    table = np.zeros(256, dtype=np.uint8)
    sc = lib().ora_encoding_table(b"".join(groups), offs.ctypes.data_as(C.POINTER(C.c_uint32)), len(groups),
                                  1 if with_wildcard else 0, table.ctypes.data_as(C.POINTER(C.c_uint8)))
    return table, int(sc)


def aligned_empty(nbytes: int, align: int = 64) -> np.ndarray:
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


def blob_layout(t: IndexType, text_len: int, symbol_count: int, kmer_size: int = 1, sampling_ratio: int = 1) -> OraLayout:
    L = OraLayout()
    detail = (C.c_uint64 * 2)()
    rc = lib().ora_builder_layout(t.c(), text_len, symbol_count, kmer_size, sampling_ratio, C.byref(L), detail)
    if rc:
        raise OracleError(rc, detail)
    return L


def build_blob(t: IndexType, text: bytes | np.ndarray, symbol_count: int, table=None, kmer_size: int = 1,
               sampling_ratio: int = 1) -> np.ndarray:
    """FmIndexBuilder::new(..).set_lookup_table_config(..).set_suffix_array_config(..).build(text, blob)."""
    text = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, dtype=np.uint8)
    L = blob_layout(t, len(text), symbol_count, kmer_size, sampling_ratio)
    blob = aligned_empty(int(L.total_size))
    detail = (C.c_uint64 * 2)()
    tbl = None
    if t.encoding_table:
        tbl = np.ascontiguousarray(table, dtype=np.uint8)
        assert tbl.size == 256
    rc = lib().ora_build(t.c(), text.ctypes.data, len(text), symbol_count,
                         tbl.ctypes.data if tbl is not None else None, kmer_size, sampling_ratio,
                         blob.ctypes.data, blob.size, detail)
    if rc:
        raise OracleError(rc, detail)
    return blob


class OracleFmIndex:
    """FmIndex<'a, P, B, E> over a borrowed blob (CPU oracle)."""

    def __init__(self, blob: np.ndarray, t: IndexType):
        self._blob = blob  # keep alive: the oracle borrows it like the Rust lifetime 'a
        self.type = t
        h = C.c_void_p()
        detail = (C.c_uint64 * 2)()
        rc = lib().ora_load(blob.ctypes.data, blob.size, t.c(), C.byref(h), detail)
        if rc:
            raise OracleError(rc, detail)
        self._h = h

    @classmethod
    def load(cls, blob: np.ndarray, t: IndexType) -> "OracleFmIndex":
        return cls(blob, t)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib().ora_free(h)
            self._h = None

    @property
    def layout(self) -> OraLayout:
        return lib().ora_index_layout(self._h).contents

    @property
    def text_len(self) -> int:
        return int(lib().ora_text_len(self._h))

    def pos_range(self, pattern: bytes, reversed_: bool = False):
        sp, ep = C.c_uint64(), C.c_uint64()
        rc = lib().ora_pos_range(self._h, bytes(pattern), len(pattern), int(reversed_), C.byref(sp), C.byref(ep))
        if rc:
            raise OracleError(rc)
        return int(sp.value), int(ep.value)

    def count(self, pattern: bytes) -> int:
        c = C.c_uint64()
        rc = lib().ora_count(self._h, bytes(pattern), len(pattern), 0, C.byref(c))
        if rc:
            raise OracleError(rc)
        return int(c.value)

    def count_rev_iter(self, pattern_rev: bytes) -> int:
        c = C.c_uint64()
        rc = lib().ora_count(self._h, bytes(pattern_rev), len(pattern_rev), 1, C.byref(c))
        if rc:
            raise OracleError(rc)
        return int(c.value)

    def _locate(self, pattern: bytes, rev: int) -> np.ndarray:
        n = C.c_uint64(0)
        rc = lib().ora_locate(self._h, bytes(pattern), len(pattern), rev, None, 0, C.byref(n))
        if rc:
            raise OracleError(rc)
        out = np.empty(int(n.value), dtype=np.uint64)
        n2 = C.c_uint64(0)
        lib().ora_locate(self._h, bytes(pattern), len(pattern), rev, out.ctypes.data, out.size, C.byref(n2))
        return out

    def locate(self, pattern: bytes) -> np.ndarray:
        """SA-row order (unsorted), like FmIndex::locate."""
        return self._locate(pattern, 0)

    def locate_rev_iter(self, pattern_rev: bytes) -> np.ndarray:
        return self._locate(pattern_rev, 1)

    # --- pattern-parallel drivers (CPU baseline) ---
    def count_batch(self, pats: np.ndarray, threads: int = 1) -> np.ndarray:
        pats = np.ascontiguousarray(pats, dtype=np.uint8)
        n, ln = pats.shape
        out = np.empty(n, dtype=np.uint64)
        rc = lib().ora_count_batch(self._h, pats.ctypes.data, n, ln, out.ctypes.data, threads)
        if rc:
            raise OracleError(rc)
        return out

    def locate_batch(self, pats: np.ndarray, threads: int = 1, want_positions: bool = True):
        """Returns (counts u64[n], offsets u64[n+1], positions P[total] or None, checksum)."""
        pats = np.ascontiguousarray(pats, dtype=np.uint8)
        n, ln = pats.shape
        counts = np.empty(n, dtype=np.uint64)
        ck = C.c_uint64()
        if not want_positions:
            rc = lib().ora_locate_batch(self._h, pats.ctypes.data, n, ln, counts.ctypes.data, None, None, 0,
                                        C.byref(ck), threads)
            if rc:
                raise OracleError(rc)
            return counts, None, None, int(ck.value)
        rc = lib().ora_count_batch(self._h, pats.ctypes.data, n, ln, counts.ctypes.data, threads)
        if rc:
            raise OracleError(rc)
        offs = np.zeros(n + 1, dtype=np.uint64)
        np.cumsum(counts, out=offs[1:])
        pos = np.empty(int(offs[-1]), dtype=self.type.pos_dtype)
        rc = lib().ora_locate_batch(self._h, pats.ctypes.data, n, ln, None, offs.ctypes.data, pos.ctypes.data,
                                    self.type.pos_bits, C.byref(ck), threads)
        if rc:
            raise OracleError(rc)
        return counts, offs, pos, int(ck.value)


def suffix_array(text_with_sentinel: np.ndarray, alphabet: int) -> np.ndarray:
    s = np.ascontiguousarray(text_with_sentinel, dtype=np.uint8)
    sa = np.empty(s.size, dtype=np.int32)
    rc = lib().ora_suffix_array(s.ctypes.data, s.size, alphabet, sa.ctypes.data)
    if rc:
        raise OracleError(rc)
    return sa
