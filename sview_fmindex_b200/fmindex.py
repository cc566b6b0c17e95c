"""Host-side mirror of the reference's public API for the hot path, over the C ABI (include/svfm.h).

Names, argument meaning and error behaviour follow baku4/sview-fmindex (citations relative to its
`sview-fmindex/src/`):

  FmIndex::<P,B,E>::load(blob)            load_from_blob.rs:28     -> FmIndex.load(blob, index_type)
  count / locate / locate_to_buffer       locate/with_slice.rs:5-18
  count_rev_iter / locate_rev_iter(...)   locate/with_rev_iter.rs:5-18
  LoadError::{InvalidFormat, MismatchedBlobSize(expected, actual)}   load_from_blob.rs:16-24
  FmIndexBuilder::new / set_*_config / blob_size / build             builder/mod.rs:63-264
  BuildError::{SymbolCountOver, InvalidBlobSize, InvalidConfig, ...} builder/mod.rs:37-57
  EncodingTable::from_symbols[_with_wildcard] / symbol_count         encoding_table.rs:15-37

plus the batched entry points (count_batch / locate_batch) this engine adds.  Everything runs on the GPU:
there is no CPU fallback, and the CPU oracle under oracle/ is never imported from here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import SvfmInfo, SvfmType


# ---- errors ------------------------------------------------------------------------------------
class SvfmError(Exception):
    def __init__(self, code: int, detail=(0, 0), msg: str = ""):
        self.code = code
        self.detail = tuple(int(x) for x in detail)
        name = _ffi.ERROR_NAMES.get(code, str(code))
        super().__init__(f"{name}{' ' + msg if msg else ''} detail={self.detail}")


class LoadError(SvfmError):
    """LoadError (load_from_blob.rs:16-24)."""


class InvalidFormat(LoadError):
    pass


class MismatchedBlobSize(LoadError):
    @property
    def expected(self):
        return self.detail[0]

    @property
    def actual(self):
        return self.detail[1]


class BuildError(SvfmError):
    """BuildError (builder/mod.rs:37-57)."""


class EmptyPattern(SvfmError):
    """The reference panics on an empty pattern (count_array.rs:211); here it is an error."""


def _raise(code: int, detail=(0, 0)):
    if code == _ffi.SVFM_OK:
        return
    msg = ""
    if code == _ffi.SVFM_ERR_CUDA:
        msg = (_ffi.lib().svfm_last_error() or b"").decode(errors="replace")
    if code == _ffi.SVFM_ERR_INVALID_FORMAT:
        raise InvalidFormat(code, detail, msg)
    if code == _ffi.SVFM_ERR_BLOB_SIZE:
        raise MismatchedBlobSize(code, detail, msg)
    if code in (_ffi.SVFM_ERR_SYMBOL_COUNT_OVER, _ffi.SVFM_ERR_TEXT_LENGTH, _ffi.SVFM_ERR_INVALID_BLOB_SIZE,
                _ffi.SVFM_ERR_NOT_ALIGNED, _ffi.SVFM_ERR_INVALID_CONFIG):
        raise BuildError(code, detail, msg)
    if code == _ffi.SVFM_ERR_EMPTY_PATTERN:
        raise EmptyPattern(code, detail, msg)
    raise SvfmError(code, detail, msg)


# ---- the (P, B, E) triple ------------------------------------------------------------------------
@dataclass(frozen=True)
class IndexType:
    """FmIndex<'a, P, B, E>: IndexType(32, 3, 64, True) = <u32, Block3<u64>, EncodingTable>."""
    pos_bits: int = 32
    planes: int = 3
    vec_bits: int = 64
    encoding_table: bool = True

    def c(self) -> SvfmType:
        return SvfmType(self.pos_bits, self.planes, self.vec_bits, 1 if self.encoding_table else 0)

    @property
    def pos_dtype(self):
        return np.uint32 if self.pos_bits == 32 else np.uint64

    def __str__(self):
        return f"<u{self.pos_bits}, Block{self.planes}<u{self.vec_bits}>, {'EncodingTable' if self.encoding_table else 'PassThrough'}>"


class EncodingTable:
    """EncodingTable([u8; 256]) (encoding_table.rs:7-38): byte -> symbol index, unmapped bytes -> last symbol."""

    def __init__(self, table: np.ndarray):
        self.table = np.ascontiguousarray(table, dtype=np.uint8)
        assert self.table.size == 256

    @classmethod
    def from_symbols(cls, symbols) -> "EncodingTable":
        groups = [bytes(s) for s in symbols]
        t = np.full(256, len(groups) - 1, dtype=np.uint8)
        for i, g in enumerate(groups):
            for b in g:
                t[b] = i
        return cls(t)

    @classmethod
    def from_symbols_with_wildcard(cls, symbols) -> "EncodingTable":
        groups = [bytes(s) for s in symbols]
        t = np.full(256, len(groups), dtype=np.uint8)
        for i, g in enumerate(groups):
            for b in g:
                t[b] = i
        return cls(t)

    def symbol_count(self) -> int:
        return int(self.table.max()) + 1

    def idx_of(self, sym: int) -> int:
        return int(self.table[sym])


def _as_u8(buf) -> np.ndarray:
    if isinstance(buf, np.ndarray):
        return np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def _pack_patterns(patterns):
    """-> (bytes u8[], offs u64[n+1] or None, n, fixed_len).  A 2-D uint8 array is a fixed-length batch; a
    (bytes u8[], offs u64[n+1]) tuple is taken as is (the C ABI's own form)."""
    if isinstance(patterns, tuple) and len(patterns) == 2 and isinstance(patterns[1], np.ndarray):
        data = np.ascontiguousarray(patterns[0], dtype=np.uint8).reshape(-1)
        offs = np.ascontiguousarray(patterns[1], dtype=np.uint64)
        return data, offs, len(offs) - 1, 0
    if isinstance(patterns, np.ndarray) and patterns.ndim == 2:
        p = np.ascontiguousarray(patterns, dtype=np.uint8)
        return p.reshape(-1), None, p.shape[0], p.shape[1]
    pats = [bytes(p) for p in patterns]
    offs = np.zeros(len(pats) + 1, dtype=np.uint64)
    if pats:
        offs[1:] = np.cumsum([len(p) for p in pats], dtype=np.uint64)
    data = np.frombuffer(b"".join(pats), dtype=np.uint8) if pats else np.zeros(0, dtype=np.uint8)
    return data, offs, len(pats), 0


class FmIndex:
    """Device-resident FmIndex<'a, P, B, E>.  Owns a byte-for-byte copy of the blob in HBM."""

    def __init__(self, handle, index_type: IndexType, blob=None):
        self._h = handle
        self.type = index_type
        self._blob = blob  # kept for blob() like the Rust view keeps &'a [u8] (reference_to_source_blob.rs:9)

    @classmethod
    def load(cls, blob, index_type: IndexType, device: int = 0) -> "FmIndex":
        b = _as_u8(blob)
        h = C.c_void_p()
        detail = (C.c_uint64 * 2)()
        rc = _ffi.lib().svfm_load(b.ctypes.data, b.size, index_type.c(), device, C.byref(h), detail)
        _raise(rc, detail)
        return cls(h, index_type, b)

    @classmethod
    def load_device(cls, d_ptr: int, nbytes: int, index_type: IndexType, device: int = 0) -> "FmIndex":
        h = C.c_void_p()
        detail = (C.c_uint64 * 2)()
        rc = _ffi.lib().svfm_load_device(d_ptr, nbytes, index_type.c(), device, C.byref(h), detail)
        _raise(rc, detail)
        return cls(h, index_type, None)

    @staticmethod
    def check_blob(blob, index_type: IndexType) -> SvfmInfo:
        """Host-only validation (the LoadError paths of FmIndex::load) -- needs no device."""
        b = _as_u8(blob)
        info = SvfmInfo()
        detail = (C.c_uint64 * 2)()
        rc = _ffi.lib().svfm_check_blob(b.ctypes.data, b.size, index_type.c(), C.byref(info), detail)
        _raise(rc, detail)
        return info

    def close(self):
        if getattr(self, "_h", None):
            _ffi.lib().svfm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def blob(self):
        return self._blob

    @property
    def handle(self):
        return self._h

    def info(self) -> SvfmInfo:
        info = SvfmInfo()
        _raise(_ffi.lib().svfm_index_info(self._h, C.byref(info)))
        return info

    def memory(self) -> dict:
        """Device bytes held by the handle: the blob copy, the derived structures, idle scratch."""
        out = (C.c_uint64 * 8)()
        _raise(_ffi.lib().svfm_index_memory(self._h, out))
        return {"blob": int(out[0]), "ext_table": int(out[1]), "interleaved_occ": int(out[2]), "scratch": int(out[3]),
                "text_copy": int(out[4]), "expanded_sa": int(out[5]), "sweep_occ": int(out[6])}

    # ---- single pattern: the reference's API -----------------------------------------------------
    def count(self, pattern) -> int:
        p = _as_u8(pattern)
        c = C.c_uint64()
        _raise(_ffi.lib().svfm_count(self._h, p.ctypes.data, p.size, 0, C.byref(c)))
        return int(c.value)

    def count_rev_iter(self, pattern_rev_iter) -> int:
        p = _as_u8(bytes(pattern_rev_iter))
        c = C.c_uint64()
        _raise(_ffi.lib().svfm_count(self._h, p.ctypes.data, p.size, _ffi.SVFM_REVERSED, C.byref(c)))
        return int(c.value)

    def _locate1(self, pattern, flags) -> np.ndarray:
        _, positions = self._locate_batch_raw(_as_u8(pattern).reshape(1, -1), flags)
        return positions

    def locate(self, pattern) -> np.ndarray:
        """Vec<P> in SA-row order ("The locations may not be in order", README.md:77)."""
        return self._locate1(pattern, 0)

    def locate_rev_iter(self, pattern_rev_iter) -> np.ndarray:
        return self._locate1(bytes(pattern_rev_iter), _ffi.SVFM_REVERSED)

    def locate_to_buffer(self, pattern, buffer: list):
        """Appends without clearing, like locate_to_buffer(&self, pattern, &mut Vec<P>)."""
        buffer.extend(int(x) for x in self.locate(pattern))

    def locate_rev_iter_to_buffer(self, pattern_rev_iter, buffer: list):
        buffer.extend(int(x) for x in self.locate_rev_iter(pattern_rev_iter))

    # ---- batches ----------------------------------------------------------------------------------
    def count_batch(self, patterns, reversed_: bool = False) -> np.ndarray:
        data, offs, n, fixed = _pack_patterns(patterns)
        out = np.zeros(n, dtype=self.type.pos_dtype)
        rc = _ffi.lib().svfm_count_batch(self._h, data.ctypes.data, offs.ctypes.data if offs is not None else None,
                                         n, fixed, _ffi.SVFM_REVERSED if reversed_ else 0, out.ctypes.data)
        _raise(rc)
        return out

    def _locate_batch_raw(self, patterns, flags):
        data, offs, n, fixed = _pack_patterns(patterns)
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        ptr = C.c_void_p()
        total = C.c_uint64()
        rc = _ffi.lib().svfm_locate_batch_alloc(self._h, data.ctypes.data, offs.ctypes.data if offs is not None else None,
                                                n, fixed, flags, out_offs.ctypes.data, C.byref(ptr), C.byref(total))
        _raise(rc)
        t = int(total.value)
        if t:
            pos = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32 if self.type.pos_bits == 32 else C.c_uint64)),
                                        shape=(t,)).copy()
            _ffi.lib().svfm_free_positions(ptr)
        else:
            pos = np.zeros(0, dtype=self.type.pos_dtype)
        return out_offs, pos

    def locate_batch(self, patterns, sorted_: bool = False, reversed_: bool = False, offs32: bool = False):
        """-> (out_offs u64[n+1] (u32 with offs32), positions P[total]); pattern i owns positions[out_offs[i]:out_offs[i+1]]."""
        flags = (_ffi.SVFM_SORTED if sorted_ else 0) | (_ffi.SVFM_REVERSED if reversed_ else 0)
        if not offs32:
            return self._locate_batch_raw(patterns, flags)
        data, offs, n, fixed = _pack_patterns(patterns)
        out_offs = np.zeros(n + 1, dtype=np.uint32)
        ptr, total = C.c_void_p(), C.c_uint64()
        _raise(_ffi.lib().svfm_locate_batch_alloc(self._h, data.ctypes.data, offs.ctypes.data if offs is not None else None, n, fixed,
                                                  flags | _ffi.SVFM_OFFS32, out_offs.ctypes.data, C.byref(ptr), C.byref(total)))
        return out_offs, self._take_positions(ptr, int(total.value))

    def _take_positions(self, ptr, t: int) -> np.ndarray:
        if not t:
            return np.zeros(0, dtype=self.type.pos_dtype)
        pos = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32 if self.type.pos_bits == 32 else C.c_uint64)), shape=(t,)).copy()
        _ffi.lib().svfm_free_positions(ptr)
        return pos

    # ---- packed fixed-length batches (include/svfm.h, "packed fixed-length batches") ----------------------------
    @staticmethod
    def pack_patterns(patterns: np.ndarray, table, bits: int) -> np.ndarray:
        """2-D uint8 patterns -> packed (n, ceil(len*bits/8)) uint8; table = 256-entry encoding table or None."""
        p = np.ascontiguousarray(patterns, dtype=np.uint8)
        n, ln = p.shape
        out = np.zeros((n, (ln * bits + 7) // 8), dtype=np.uint8)
        tbl = np.ascontiguousarray(table, dtype=np.uint8) if table is not None else None
        _raise(_ffi.lib().svfm_pack_patterns(p.ctypes.data, n, ln, tbl.ctypes.data if tbl is not None else None, bits, out.ctypes.data))
        return out

    def count_batch_packed(self, packed: np.ndarray, length: int, bits: int) -> np.ndarray:
        pk = np.ascontiguousarray(packed, dtype=np.uint8)
        n = pk.shape[0]
        out = np.zeros(n, dtype=self.type.pos_dtype)
        _raise(_ffi.lib().svfm_count_batch_packed(self._h, pk.ctypes.data, n, length, bits, 0, out.ctypes.data))
        return out

    def locate_batch_packed(self, packed: np.ndarray, length: int, bits: int, offs32: bool = False, sorted_: bool = False):
        pk = np.ascontiguousarray(packed, dtype=np.uint8)
        n = pk.shape[0]
        flags = (_ffi.SVFM_SORTED if sorted_ else 0) | (_ffi.SVFM_OFFS32 if offs32 else 0)
        out_offs = np.zeros(n + 1, dtype=np.uint32 if offs32 else np.uint64)
        total = C.c_uint64()
        rc = _ffi.lib().svfm_locate_batch_packed(self._h, pk.ctypes.data, n, length, bits, flags, out_offs.ctypes.data, None, 0, C.byref(total))
        if rc not in (_ffi.SVFM_OK, _ffi.SVFM_ERR_CAPACITY):
            _raise(rc)
        pos = np.zeros(int(total.value), dtype=self.type.pos_dtype)
        if pos.size:
            _raise(_ffi.lib().svfm_locate_batch_packed(self._h, pk.ctypes.data, n, length, bits, flags, out_offs.ctypes.data,
                                                       pos.ctypes.data, pos.size, C.byref(total)))
        return out_offs, pos


# ---- builder (GPU suffix sort) ----------------------------------------------
class SuffixArrayConfig:
    """build_config::SuffixArrayConfig (suffix_array_config.rs:5-34)."""
    Uncompressed = ("Uncompressed", 1)

    @staticmethod
    def Compressed(ratio: int):
        return ("Compressed", int(ratio))


class LookupTableConfig:
    """build_config::LookupTableConfig (lookup_table_config.rs:6-53)."""
    NONE = ("None", 1)

    @staticmethod
    def KmerSize(k: int):
        return ("KmerSize", int(k))

    @staticmethod
    def MaxMemory(nbytes: int):
        return ("MaxMemory", int(nbytes))


class FmIndexBuilder:
    """FmIndexBuilder<P, B, E> (builder/mod.rs:18-264); build() runs the suffix sort on the GPU."""

    def __init__(self, text_len: int, symbol_count: int, text_encoder, index_type: IndexType):
        self.text_len = int(text_len)
        self.symbol_count = int(symbol_count)
        self.text_encoder = text_encoder  # EncodingTable or None (PassThrough)
        self.type = index_type
        if (text_encoder is not None) != index_type.encoding_table:
            raise SvfmError(_ffi.SVFM_ERR_BAD_TYPE, msg="encoder does not match the index type")
        self.kmer_size = 1
        self.sampling_ratio = 1
        self.blob_size()  # validates like FmIndexBuilder::new (SymbolCountOver)

    @classmethod
    def new(cls, text_len, symbol_count, text_encoder, index_type) -> "FmIndexBuilder":
        return cls(text_len, symbol_count, text_encoder, index_type)

    def set_lookup_table_config(self, config) -> "FmIndexBuilder":
        kind, v = config
        if kind == "KmerSize":
            if v < 2:
                raise BuildError(_ffi.SVFM_ERR_INVALID_CONFIG, msg="K-mer size must be at least 2")
            self.kmer_size = v
        elif kind == "MaxMemory":
            swsc, k = self.symbol_count + 1, 2
            # lookup_table_config.rs:41-53, plus: (S+1)^k must fit the u32 the reference stores it in (count_array.rs:70)
            while swsc ** k * (self.type.pos_bits // 8) <= v and swsc ** k <= 0xFFFFFFFF:
                k += 1
            self.kmer_size = k - 1
        else:
            self.kmer_size = 1
        self.blob_size()
        return self

    def set_suffix_array_config(self, config) -> "FmIndexBuilder":
        kind, v = config
        if kind == "Compressed":
            if v < 2:
                raise BuildError(_ffi.SVFM_ERR_INVALID_CONFIG,
                                 msg="Sampling ratio for compressed suffix array must be at least 2")
            self.sampling_ratio = v
        else:
            self.sampling_ratio = 1
        self.blob_size()
        return self

    def blob_size(self) -> int:
        size = C.c_uint64()
        detail = (C.c_uint64 * 2)()
        rc = _ffi.lib().svfm_blob_size(self.type.c(), self.text_len, self.symbol_count, self.kmer_size,
                                       self.sampling_ratio, C.byref(size), detail)
        _raise(rc, detail)
        return int(size.value)

    def build(self, text, blob: np.ndarray, device: int = 0):
        t = _as_u8(text)
        if t.size != self.text_len:
            raise BuildError(_ffi.SVFM_ERR_TEXT_LENGTH, (self.text_len, t.size))
        detail = (C.c_uint64 * 2)()
        tbl = self.text_encoder.table.ctypes.data if self.text_encoder is not None else None
        rc = _ffi.lib().svfm_build(self.type.c(), t.ctypes.data, t.size, self.symbol_count, tbl, self.kmer_size,
                                   self.sampling_ratio, device, blob.ctypes.data, blob.size, detail)
        _raise(rc, detail)

    def build_device(self, d_text: int, d_blob: int, blob_len: int, device: int = 0):
        detail = (C.c_uint64 * 2)()
        tbl = self.text_encoder.table.ctypes.data if self.text_encoder is not None else None
        rc = _ffi.lib().svfm_build_device(self.type.c(), d_text, self.text_len, self.symbol_count, tbl,
                                          self.kmer_size, self.sampling_ratio, device, d_blob, blob_len, detail)
        _raise(rc, detail)


def aligned_empty(nbytes: int, align: int = 64) -> np.ndarray:
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]
