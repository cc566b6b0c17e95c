"""(test infrastructure: imports the oracle as the checker)  One compact workload for compute-sanitizer (memcheck / racecheck / synccheck): every hand-written
kernel of the hot path on a small index, results checked against the CPU oracle.

    compute-sanitizer --tool memcheck  python tests/sanitize_case.py
    compute-sanitizer --tool racecheck python tests/sanitize_case.py
    compute-sanitizer --tool synccheck python tests/sanitize_case.py

Covers: GPU index builder, extended-table / interleaved-copy / text-copy construction, pack_sweep_kernel (TMA bulk copies +
mbarriers), sweep_round_kernel in all three partition modes (look-back descriptors, atomic reservation), locate in bucket and
CSR mode (incl. the heavy list), sb_scan / sb_place, scatter_counts, the generic search kernel with text verification, the
small-batch kernel, packed input, the radix sort-back; with and without the expanded suffix array and the sweep occ copy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sview_fmindex_b200 as fm  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (checker only)
from sview_fmindex_b200 import _ffi  # noqa: E402

L = _ffi.lib()
rng = np.random.default_rng(7)
n = 120_000
text = np.concatenate([np.full(3000, ord("A"), dtype=np.uint8), np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]])
n = len(text)
symbols = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]
table, sc = po.encoding_table(symbols)


def check(gpu, ora, pats, tag):
    oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
    assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc), tag
    offs, pos = gpu.locate_batch(pats)
    assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64)), tag
    print("ok", tag, len(pats), int(oo[-1]), flush=True)


for (p, planes, v, k, r) in ((32, 3, 64, 3, 2), (64, 2, 128, 2, 3)):
    it = fm.IndexType(p, planes, v, True)
    enc = fm.EncodingTable.from_symbols(symbols if planes > 2 else symbols[:4])
    b = fm.FmIndexBuilder(n, enc.symbol_count(), enc, it)
    b.kmer_size, b.sampling_ratio = k, r
    blob = fm.aligned_empty(b.blob_size())
    b.build(text, blob)                                                    # GPU builder
    ora = po.OracleFmIndex.load(blob, po.IndexType(p, planes, v, True))
    for sweep_min, bucket, ext_bits, derived in ((0, 1, 16, 1), (0, 0, 12, 0), (2**64 - 1, 1, 16, 1), (2**64 - 1, 1, 16, 0)):
        L.svfm_set_tuning(_ffi.SVFM_TUNE_FULL_SA, derived)       # 0: locate LF-walks (locate_warp_kernel), 1: locate_direct_kernel
        L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_OCC, derived)     # sweep rounds on the 32-byte occ copy / on the blob in place
        L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_MIN, sweep_min)
        L.svfm_set_tuning(_ffi.SVFM_TUNE_SORT_MIN, 0 if sweep_min == 0 else 2**64 - 1)
        L.svfm_set_tuning(_ffi.SVFM_TUNE_BUCKET_SORTBACK, bucket)
        L.svfm_set_tuning(_ffi.SVFM_TUNE_EXT_BITS, ext_bits)
        gpu = fm.FmIndex.load(blob, it)
        for ln, m in ((12, 20_000), (20, 9_000), (3, 700), (150, 3_000)):
            starts = rng.integers(0, n - ln, size=m)
            pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
            pats[::5, ln // 2] = ord("G")
            pats[1::97] = ord("A")                                         # thousands of rows each: the heavy list
            check(gpu, ora, pats, (p, planes, v, sweep_min, bucket, ln))
        var = [bytes(text[s:s + 1 + (i % 40)]) for i, s in enumerate(rng.integers(0, n - 41, size=5000))]
        offs, pos = gpu.locate_batch(var)
        for i in range(0, 5000, 611):
            assert np.array_equal(pos[int(offs[i]):int(offs[i + 1])].astype(np.uint64), ora.locate(var[i]))
        for q in (b"ACGT", b"A", b"GATTACAGATTACA", bytes(text[5000:5150])):       # single-pattern calls: small-batch kernel
            assert gpu.count(q) == ora.count(q)
            assert np.array_equal(gpu.locate(q).astype(np.uint64), ora.locate(q))
        pk = gpu.pack_patterns(pats[:2000], enc.table, 3)
        o3, p3 = gpu.locate_batch_packed(pk, 150, 3, offs32=True)
        o4, p4 = gpu.locate_batch(pats[:2000])
        assert np.array_equal(o3.astype(np.uint64), o4) and np.array_equal(p3, p4)
        gpu.close()
L.svfm_set_tuning(_ffi.SVFM_TUNE_FULL_SA, 1)
L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_OCC, 1)
print("sanitize_case: all results match the oracle")
