"""GPU parity tests: the CUDA path (through the C ABI, sview_fmindex_b200/libsvfm.so) against the CPU
oracle on identical blobs and patterns.  Bit-exact: counts, CSR offsets and positions in the reference's
SA-row order (locate/mod.rs:19).  Mirrors the reference's tests."""
import json
import os

import numpy as np
import pytest

from helpers import ALL_TYPES, brute_force_locate, gen_rand_chr_list, gen_rand_pattern, gen_rand_text

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "readme_example.json")


@pytest.fixture(scope="module", params=["reorder_always", "reorder_never", "defaults", "no_ext_table", "radix_sortback"])
def fm(request):
    """Every parity test runs with the batch reordering (sweep search for fixed-length batches, locality sort for the
    rest) forced on, forced off (and the kernels reading the blob's occ sections in place instead of the interleaved
    copy), at its default thresholds, and without the extended k-mer table (so that long patterns seed from the blob's
    own kLTS), and with the reordered batches radix-sorted back into the caller's order instead of the bucketed sort-back;
    two of the five run without the packed text copy (no text verification), one sorts the sweep items with this library's
    own radix pass instead of cub; two run without the expanded suffix array (locate LF-walks to a sampled row as the
    reference does), three with it: results must not depend on any of it."""
    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi
    L = _ffi.lib()
    never = 2**64 - 1
    sort_min, sweep_min, ext_bits, ilv, bucket, text, full_sa = {
        "reorder_always": (0, 0, 24, 1, 1, 1, 1), "reorder_never": (never, never, 24, 0, 1, 0, 0),
        "defaults": (_ffi.SVFM_TUNE_AUTO, _ffi.SVFM_TUNE_AUTO, _ffi.SVFM_TUNE_AUTO, 1, 1, 1, 1),
        "no_ext_table": (_ffi.SVFM_TUNE_AUTO, 0, 0, 1, 1, 1, 0), "radix_sortback": (0, 0, 24, 1, 0, 0, 1)}[request.param]
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_SORT_MIN, sort_min) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_MIN, sweep_min) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_EXT_BITS, ext_bits) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_ILV, ilv) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_BUCKET_SORTBACK, bucket) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_TEXT, text) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_FULL_SA, full_sa) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_OCC, 0 if request.param == "radix_sortback" else 1) == 0
    assert L.svfm_set_tuning(_ffi.SVFM_TUNE_OWN_RADIX, 1 if request.param == "radix_sortback" else 0) == 0
    yield fm
    L.svfm_set_tuning(_ffi.SVFM_TUNE_SORT_MIN, _ffi.SVFM_TUNE_AUTO)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_MIN, _ffi.SVFM_TUNE_AUTO)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_EXT_BITS, _ffi.SVFM_TUNE_AUTO)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_ILV, 1)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_BUCKET_SORTBACK, 1)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_TEXT, 1)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_FULL_SA, 1)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_SWEEP_OCC, 1)
    L.svfm_set_tuning(_ffi.SVFM_TUNE_OWN_RADIX, 0)


def _pair(po, fm, text, symbols, p, n, v, k, r, passthrough=False, with_wildcard=False):
    table, sc = po.encoding_table(symbols, with_wildcard)
    if passthrough:
        t = po.IndexType(p, n, v, False)
        blob = po.build_blob(t, table[np.frombuffer(bytes(text), dtype=np.uint8)], sc, None, k, r)
    else:
        t = po.IndexType(p, n, v, True)
        blob = po.build_blob(t, text, sc, table, k, r)
    ora = po.OracleFmIndex.load(blob, t)
    gpu = fm.FmIndex.load(blob, fm.IndexType(p, n, v, not passthrough))
    return ora, gpu, table, sc


def _check_batch(ora, gpu, patterns, reversed_=False):
    counts = gpu.count_batch(patterns, reversed_=reversed_)
    offs, pos = gpu.locate_batch(patterns, reversed_=reversed_)
    offs_s, pos_s = gpu.locate_batch(patterns, sorted_=True, reversed_=reversed_)
    assert np.array_equal(offs, offs_s)
    for i, pat in enumerate(patterns):
        pat = bytes(pat)
        exp = ora.locate_rev_iter(pat) if reversed_ else ora.locate(pat)
        assert int(counts[i]) == len(exp), (i, pat)
        got = pos[int(offs[i]):int(offs[i + 1])].astype(np.uint64)
        assert np.array_equal(got, exp), (i, pat)  # SA-row order, bit-exact
        got_s = pos_s[int(offs[i]):int(offs[i + 1])].astype(np.uint64)
        assert np.array_equal(got_s, np.sort(exp)), (i, pat)


def test_readme_golden(oracle, fm):
    g = json.load(open(GOLDEN))
    ora, gpu, _, sc = _pair(oracle, fm, g["text"].encode(), [s.encode() for s in g["symbols"]], 32, 2, 64, 1, 1)
    assert sc == 4
    for q in g["queries"]:
        pat = q["pattern"].encode()
        assert gpu.count(pat) == q["count"]
        assert sorted(int(x) for x in gpu.locate(pat)) == q["locate_sorted"]
        assert gpu.count_rev_iter(pat[::-1]) == q["count"]
        assert sorted(int(x) for x in gpu.locate_rev_iter(pat[::-1])) == q["locate_sorted"]
        buf = [7]
        gpu.locate_to_buffer(pat, buf)  # appends without clearing
        assert buf[0] == 7 and sorted(buf[1:]) == q["locate_sorted"]
    info = gpu.info()
    assert info.text_len == 31 and info.sentinel_index == g["sentinel_index"] and info.header_size == 328


@pytest.mark.parametrize("chr_count", [3, 4, 5, 7, 8, 9, 15, 16, 17, 33, 64])
def test_results_are_accurate(oracle, fm, chr_count):
    """get_accurate_result/mod.rs:61-142 on the GPU: every (P, Block, Vector) that can hold the alphabet,
    ltks=3, sasr=2; answers = brute force AND the oracle (bit-exact, SA-row order)."""
    rng = np.random.default_rng(2000 + chr_count)
    chr_list = gen_rand_chr_list(rng, chr_count)
    text = gen_rand_text(rng, chr_list, 100, 300)
    patterns = [gen_rand_pattern(rng, text, 1, 10) for _ in range(100)]
    symbols = [bytes([c]) for c in chr_list]
    for (p, n, v) in ALL_TYPES:
        if (1 << n) < chr_count:
            continue
        ora, gpu, table, _ = _pair(oracle, fm, text, symbols, p, n, v, 3, 2)
        enc_text = table[np.frombuffer(text, dtype=np.uint8)]
        offs, pos = gpu.locate_batch(patterns, sorted_=True)
        for i, pat in enumerate(patterns):
            ans = brute_force_locate(enc_text, table[np.frombuffer(pat, dtype=np.uint8)])
            assert np.array_equal(pos[int(offs[i]):int(offs[i + 1])].astype(np.uint64), ans), (p, n, v, pat)
        _check_batch(ora, gpu, patterns[:40])
        gpu.close()


def test_fixed_length_batches_every_type(oracle, fm):
    """Fixed-length batches (the shape the sweep search takes) for every (P, Block, Vector): lengths below, at and
    above the extended table's k, up to the longest pattern whose symbols fit the 64-bit item, and one beyond it;
    patterns cut from the text, mutated, absent, and with the wildcard symbol."""
    rng = np.random.default_rng(77)
    for chr_count, with_wildcard in ((4, False), (4, True), (20, True), (50, False)):
        chr_list = gen_rand_chr_list(rng, chr_count)
        n = 6000
        text = np.frombuffer(bytes(chr_list), dtype=np.uint8)[rng.integers(0, chr_count, size=n)].copy()
        symbols = [bytes([c]) for c in chr_list]
        if with_wildcard:
            text[rng.integers(0, n, size=40)] = ord("~")  # not in any symbol set -> wildcard
        for (p, nn, v) in ALL_TYPES:
            if (1 << nn) < chr_count + (1 if with_wildcard else 0):
                continue
            if (p, v) not in ((32, 64), (64, 128), (64, 32)) and chr_count != 4:
                continue  # all 30 types for DNA, a spread of them for the wider alphabets
            ora, gpu, table, sc = _pair(oracle, fm, bytes(text), symbols, p, nn, v, 2, 3, with_wildcard=with_wildcard)
            bits = max(1, int(np.ceil(np.log2(sc))))
            for ln in (1, 2, 3, 4, 5, 6, 7, 11, 64 // bits, 64 // bits + 7, 40):
                m = 300
                starts = rng.integers(0, n - ln, size=m)
                pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
                pats[::5, int(rng.integers(0, ln))] = chr_list[0]     # mutate: many become absent
                pats[7::31, ln - 1] = ord("~")                        # wildcard / unmapped byte
                oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
                assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc), (p, nn, v, ln)
                offs, pos = gpu.locate_batch(pats)
                assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64)), (p, nn, v, ln)
                rc = gpu.count_batch(pats[:, ::-1], reversed_=True)
                assert np.array_equal(rc.astype(np.uint64), oc), (p, nn, v, ln)
            gpu.close()


def test_config_invariance(oracle, fm):
    """config_invariance/mod.rs:51-143: kLTS {None,2,3,4} x SA {Uncompressed,2,3,4} x every type."""
    rng = np.random.default_rng(43)
    text = gen_rand_text(rng, b"ACGT", 1000, 1000)
    patterns = [gen_rand_pattern(rng, text, 10, 10)] + [gen_rand_pattern(rng, text, 1, 6) for _ in range(7)]
    symbols = [b"A", b"C", b"G", b"T"]
    base = None
    for k in (1, 2, 3, 4):
        for r in (1, 2, 3, 4):
            for (p, n, v) in ALL_TYPES:
                ora, gpu, _, _ = _pair(oracle, fm, text, symbols, p, n, v, k, r)
                offs, pos = gpu.locate_batch(patterns, sorted_=True)
                ans = [pos[int(offs[i]):int(offs[i + 1])].astype(np.uint64) for i in range(len(patterns))]
                if base is None:
                    base = ans
                    _check_batch(ora, gpu, patterns)
                for a, b in zip(ans, base):
                    assert np.array_equal(a, b), (k, r, p, n, v)
                gpu.close()


def test_text_encoders_are_consistent(oracle, fm):
    """text_encoders_consistency/mod.rs:86-106: slice == rev iter; EncodingTable == PassThrough."""
    rng = np.random.default_rng(4)
    for chr_count, types in ((4, [(32, 2, 64), (64, 3, 32)]), (21, [(32, 5, 64), (64, 6, 128)])):
        chr_list = gen_rand_chr_list(rng, chr_count)
        text = gen_rand_text(rng, chr_list, 300, 500)
        symbols = [bytes([c]) for c in chr_list]
        patterns = [gen_rand_pattern(rng, text, 1, 12) for _ in range(64)]
        for (p, n, v) in types:
            ora_t, gpu_t, table, _ = _pair(oracle, fm, text, symbols, p, n, v, 3, 2)
            ora_p, gpu_p, _, _ = _pair(oracle, fm, text, symbols, p, n, v, 3, 2, passthrough=True)
            enc = [bytes(table[np.frombuffer(q, dtype=np.uint8)]) for q in patterns]
            _check_batch(ora_t, gpu_t, patterns)
            _check_batch(ora_t, gpu_t, [q[::-1] for q in patterns], reversed_=True)
            _check_batch(ora_p, gpu_p, enc)
            _check_batch(ora_p, gpu_p, [q[::-1] for q in enc], reversed_=True)
            ot, pt = gpu_t.locate_batch(patterns)
            op, pp = gpu_p.locate_batch(enc)
            assert np.array_equal(ot, op) and np.array_equal(pt, pp)
            assert np.array_equal(gpu_t.count_batch(patterns), gpu_t.count_batch([q[::-1] for q in patterns], reversed_=True))


@pytest.mark.parametrize("vec_bits", [32, 64, 128])
def test_edge_cases(oracle, fm, vec_bits):
    """SURVEY.md Appendix B on the GPU."""
    rng = np.random.default_rng(12 + vec_bits)
    symbols = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]
    for n in (vec_bits * 4, vec_bits, vec_bits - 1, vec_bits + 1, 1, 2, 3):
        text = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])
        for k, r in ((1, 1), (3, 2), (4, 5)):
            ora, gpu, _, _ = _pair(oracle, fm, text, symbols, 32, 3, vec_bits, k, r)
            pats = {text[i:i + m] for i in range(0, n, max(1, n // 7)) for m in (1, 2, 3, 4, 9) if i + m <= n}
            pats |= {b"A", b"AC", b"ACG", b"TTTTTTTT", b"N", b"NN", b"xyz", b"GATTACA", b"a", b"acgt"}
            _check_batch(ora, gpu, sorted(pats))
            gpu.close()
    # homopolymer: one pattern with hundreds of rows, LF walks that hit the sentinel row
    ora, gpu, _, _ = _pair(oracle, fm, b"A" * 777, [b"A", b"C", b"G", b"T"], 64, 2, vec_bits, 2, 3)
    _check_batch(ora, gpu, [b"AAAA", b"A", b"C", b"AC", b"A" * 777, b"A" * 778])
    assert gpu.count(b"AAAA") == 774
    # more than HEAVY_ROWS (1024) rows per pattern: the row-parallel locate kernel, mixed with light patterns
    text = b"A" * 3000 + b"CGT" * 700 + b"A" * 2500
    ora_h, gpu_h, _, _ = _pair(oracle, fm, text, [b"A", b"C", b"G", b"T"], 32, 2, vec_bits, 2, 4)
    _check_batch(ora_h, gpu_h, [b"A", b"AA", b"CGT", b"GTC", b"T", b"AAAAAAAAAAAAAAAAAAAAAAAA", b"TA", b"CC", b"ACGT"] * 5)
    # empty pattern: reference panics (count_array.rs:211) -> error code, nothing computed
    with pytest.raises(fm.EmptyPattern):
        gpu.count(b"")
    with pytest.raises(fm.EmptyPattern):
        gpu.count_batch([b"A", b"", b"C"])
    with pytest.raises(fm.EmptyPattern):
        gpu.locate_batch([b"A", b""])
    # empty batch
    assert gpu.count_batch([]).size == 0
    offs, pos = gpu.locate_batch([])
    assert list(offs) == [0] and pos.size == 0
    # PassThrough byte >= symbol_count: caller error in the reference (silent wrong row); an error here
    ora_p, gpu_p, _, _ = _pair(oracle, fm, b"ACGTACGT", [b"A", b"C", b"G", b"T"], 32, 2, vec_bits, 1, 1, passthrough=True)
    with pytest.raises(fm.SvfmError) as e:
        gpu_p.count_batch([bytes([0, 1]), bytes([9, 1])])
    assert e.value.code == 24


def test_load_errors(oracle, fm):
    po = oracle
    table, sc = po.encoding_table([b"A", b"C", b"G", b"T"])
    t = po.IndexType(32, 2, 64, True)
    ft = fm.IndexType(32, 2, 64, True)
    blob = po.build_blob(t, b"ACGTACGTTTGACCA", sc, table, 2, 2)
    bad = blob.copy()
    bad[1] = ord("x")
    with pytest.raises(fm.InvalidFormat):
        fm.FmIndex.load(bad, ft)
    longer = np.concatenate([blob, np.zeros(8, dtype=np.uint8)])
    with pytest.raises(fm.MismatchedBlobSize) as e:
        fm.FmIndex.load(longer, ft)
    assert (e.value.expected, e.value.actual) == (blob.size, blob.size + 8)
    with pytest.raises(fm.MismatchedBlobSize):
        fm.FmIndex.load(blob, fm.IndexType(64, 2, 64, True))
    with pytest.raises(fm.InvalidFormat):
        fm.FmIndex.load(blob[:64], ft)
    ix = fm.FmIndex.load(blob, ft)
    assert ix.count(b"ACG") == 2
    mem = ix.memory()  # the blob copy byte for byte; derived structures only when enabled
    assert mem["blob"] == blob.size and set(mem) == {"blob", "ext_table", "interleaved_occ", "scratch", "text_copy", "expanded_sa", "sweep_occ"}


def test_medium_random_batch(oracle, fm):
    """cfg1-shaped index at 2 Mbp: u32, Block3<u64>, S=5, r=2, k=3, 20 bp patterns cut from the text plus
    absent / short / long / wildcard patterns, compared with the oracle's pattern-parallel driver."""
    po = oracle
    rng = np.random.default_rng(42)
    n = 2_000_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
    text = text.copy()
    text[rng.integers(0, n, size=2000)] = ord("N")
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    t = po.IndexType(32, 3, 64, True)
    blob = po.build_blob(t, text, sc, table, 3, 2)
    ora = po.OracleFmIndex.load(blob, t)
    gpu = fm.FmIndex.load(blob, fm.IndexType(32, 3, 64, True))
    for ln, m in ((20, 200_000), (6, 50_000), (1, 256), (2, 256), (150, 20_000)):
        starts = rng.integers(0, n - ln, size=m)
        pats = text[starts[:, None] + np.arange(ln)[None, :]]
        if ln == 20:
            pats[::7, 3] = ord("A")  # mutate some patterns so that many are absent (early exit path)
        ocounts, ooffs, opos, ock = ora.locate_batch(pats, threads=os.cpu_count() or 1)
        counts = gpu.count_batch(pats)
        offs, pos = gpu.locate_batch(pats)
        assert np.array_equal(counts.astype(np.uint64), ocounts)
        assert np.array_equal(offs, ooffs)
        assert np.array_equal(pos, opos)
        offs_s, pos_s = gpu.locate_batch(pats, sorted_=True)
        assert np.array_equal(offs_s, ooffs)
        idx = np.repeat(np.arange(m, dtype=np.uint64), np.diff(ooffs).astype(np.int64))
        exp_sorted = opos[np.lexsort((opos, idx))]
        assert np.array_equal(pos_s, exp_sorted)
        ck = int((((pos.astype(np.uint64) + np.uint64(1)) * (np.uint64(2) * idx + np.uint64(1)))).sum(dtype=np.uint64))
        assert ck == ock
    # u64 positions + Block2<u128> + sampling ratio 16 + non-power-of-two ratio
    for (p, nn, v, k, r) in ((64, 2, 128, 2, 16), (64, 3, 64, 3, 5), (32, 6, 32, 1, 7)):
        sub = text[:300_000]
        sub = np.where(sub == ord("N"), ord("A"), sub).astype(np.uint8)
        tb, s2 = po.encoding_table([b"A", b"C", b"G", b"T"])
        tt = po.IndexType(p, nn, v, True)
        b2 = po.build_blob(tt, sub, s2, tb, k, r)
        o2 = po.OracleFmIndex.load(b2, tt)
        g2 = fm.FmIndex.load(b2, fm.IndexType(p, nn, v, True))
        starts = rng.integers(0, len(sub) - 9, size=30_000)
        pats = sub[starts[:, None] + np.arange(9)[None, :]]
        oc, oo, op_, _ = o2.locate_batch(pats, threads=os.cpu_count() or 1)
        offs, pos = g2.locate_batch(pats)
        assert np.array_equal(offs, oo) and np.array_equal(pos, op_)
        assert np.array_equal(g2.count_batch(pats).astype(np.uint64), oc)


def test_chunked_host_pipeline(oracle, fm):
    """The host-buffer entry points cut big batches into chunks that flow through several streams / host
    threads: results (CSR offsets with per-chunk bases, positions, counts, capacity handling) must not change."""
    import ctypes as C
    from sview_fmindex_b200 import _ffi
    po = oracle
    L = _ffi.lib()
    rng = np.random.default_rng(99)
    n = 400_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    t = po.IndexType(32, 3, 64, True)
    blob = po.build_blob(t, text, sc, table, 3, 2)
    ora = po.OracleFmIndex.load(blob, t)
    gpu = fm.FmIndex.load(blob, fm.IndexType(32, 3, 64, True))
    m = 50_000
    starts = rng.integers(0, n - 8, size=m)
    pats = text[starts[:, None] + np.arange(8)[None, :]].copy()
    pats[::5, 2] = ord("N")
    oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
    var = [bytes(p[: 3 + (i % 6)]) for i, p in enumerate(pats[:20_000])]
    vo, vp = None, None
    try:
        for chunk in (777, 4096, 25_000, 0):
            assert L.svfm_set_tuning(_ffi.SVFM_TUNE_CHUNK, chunk) == 0
            assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc)
            offs, pos = gpu.locate_batch(pats)
            assert np.array_equal(offs, oo) and np.array_equal(pos, op_), chunk
            offs_s, pos_s = gpu.locate_batch(pats, sorted_=True)
            idx = np.repeat(np.arange(m, dtype=np.uint64), np.diff(oo).astype(np.int64))
            assert np.array_equal(offs_s, oo) and np.array_equal(pos_s, op_[np.lexsort((op_, idx))])
            # variable-length patterns through the offsets path
            o2, p2 = gpu.locate_batch(var)
            if vo is None:
                vo, vp = o2, p2
                for i in (0, 1, 7, 19_999):
                    assert np.array_equal(p2[int(o2[i]):int(o2[i + 1])].astype(np.uint64), ora.locate(var[i]))
            assert np.array_equal(o2, vo) and np.array_equal(p2, vp)
            # caller-provided buffer: exact capacity works, one short reports SVFM_ERR_CAPACITY + the needed total
            total = int(oo[-1])
            flat = np.ascontiguousarray(pats).reshape(-1)
            out_offs = np.zeros(m + 1, dtype=np.uint64)
            out_pos = np.zeros(total, dtype=np.uint32)
            tot = C.c_uint64()
            rc = L.svfm_locate_batch(gpu.handle, flat.ctypes.data, None, m, 8, 0, out_offs.ctypes.data, out_pos.ctypes.data, total, C.byref(tot))
            assert rc == 0 and tot.value == total and np.array_equal(out_pos, op_) and np.array_equal(out_offs, oo)
            rc = L.svfm_locate_batch(gpu.handle, flat.ctypes.data, None, m, 8, 0, out_offs.ctypes.data, out_pos.ctypes.data, total - 1, C.byref(tot))
            assert rc == _ffi.SVFM_ERR_CAPACITY and tot.value == total and np.array_equal(out_offs, oo)
            # an empty pattern in a late chunk fails the whole call before any work
            with pytest.raises(fm.EmptyPattern):
                gpu.locate_batch(var[:5000] + [b""] + var[5000:6000])
    finally:
        L.svfm_set_tuning(_ffi.SVFM_TUNE_CHUNK, _ffi.SVFM_TUNE_AUTO)


def test_heavy_patterns_in_fixed_length_batches(oracle, fm):
    """Patterns with thousands of SA rows (more than HEAVY_ROWS = 1024: the row-parallel locate kernel) inside
    fixed-length batches, i.e. on the sweep search when it is forced on, mixed with single-hit and absent patterns."""
    rng = np.random.default_rng(31)
    text = np.concatenate([np.full(6000, ord("A"), dtype=np.uint8),
                           np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=30_000)],
                           np.frombuffer(b"CG" * 2500, dtype=np.uint8)])
    n = len(text)
    for (p, nn, v, k, r) in ((32, 3, 64, 3, 2), (64, 2, 128, 2, 4), (32, 2, 32, 1, 1)):
        symbols = [b"Aa", b"Cc", b"Gg", b"Tt"] + ([b"Nn"] if nn > 2 else [])
        ora, gpu, _, _ = _pair(oracle, fm, bytes(text), symbols, p, nn, v, k, r)
        for ln in (2, 6, 9, 17):
            starts = rng.integers(0, n - ln, size=400)
            pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
            pats[::4] = ord("A")                                  # thousands of rows each
            pats[1::8] = np.frombuffer((b"CG" * ln)[:ln], dtype=np.uint8)
            pats[2::16, ln - 1] = ord("T")
            oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
            assert int(oc.max()) > 1024
            assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc), (p, nn, v, ln)
            offs, pos = gpu.locate_batch(pats)
            assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64)), (p, nn, v, ln)
            offs_s, pos_s = gpu.locate_batch(pats[:60], sorted_=True)
            for i in range(60):
                assert np.array_equal(pos_s[int(offs_s[i]):int(offs_s[i + 1])].astype(np.uint64), np.sort(ora.locate(bytes(pats[i]))))
        gpu.close()


def test_randomized_configs(oracle, fm):
    """Seeded fuzz over the whole configuration space: type triple, alphabet size (with / without wildcard), text length
    around the block boundaries, kLTS k, SA ratio, fixed- and variable-length batches, slice and reversed input,
    EncodingTable and PassThrough -- always bit-exact against the oracle on the same blob."""
    rng = np.random.default_rng(20261018)
    for case in range(24):
        p, nn, v = ALL_TYPES[int(rng.integers(0, len(ALL_TYPES)))]
        max_sym = 1 << nn
        with_wildcard = bool(rng.integers(0, 2))
        chr_count = int(rng.integers(2, max_sym + 1 - (1 if with_wildcard else 0))) if max_sym > 2 + with_wildcard else 2 - 0
        chr_count = max(2, min(chr_count, max_sym - (1 if with_wildcard else 0)))
        chr_list = gen_rand_chr_list(rng, chr_count)
        n = int(rng.choice([v - 1, v, v + 1, 3 * v, 1000, 4097, 20000]))
        text = np.frombuffer(bytes(chr_list), dtype=np.uint8)[rng.integers(0, chr_count, size=n)].copy()
        if with_wildcard and n > 10:
            text[rng.integers(0, n, size=max(1, n // 97))] = ord("~")
        k = int(rng.integers(1, 5))
        r = int(rng.choice([1, 2, 3, 4, 7, 16]))
        passthrough = bool(rng.integers(0, 4) == 0) and not with_wildcard
        symbols = [bytes([c]) for c in chr_list]
        ora, gpu, table, sc = _pair(oracle, fm, bytes(text), symbols, p, nn, v, k, r, passthrough=passthrough,
                                    with_wildcard=with_wildcard)
        src = table[text] if passthrough else text
        # fixed-length batch: short patterns (sweep-eligible when forced on) and, every other case, lengths up to 200
        # (beyond what a sweep item holds: plain kernel, with and without text verification / the expanded suffix array)
        ln = int(rng.integers(1, min(n, 24 if case % 2 == 0 else 200) + 1))
        m = 257
        starts = rng.integers(0, n - ln + 1, size=m)
        pats = src[starts[:, None] + np.arange(ln)[None, :]].copy()
        if ln > 1:
            pats[::3, int(rng.integers(0, ln))] = src[0]
        oc, oo, op_, _ = ora.locate_batch(pats, threads=2)
        assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc), (case, p, nn, v, chr_count, n, k, r, ln)
        offs, pos = gpu.locate_batch(pats)
        assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64)), (case, p, nn, v, n, k, r, ln)
        assert np.array_equal(gpu.count_batch(pats[:, ::-1], reversed_=True).astype(np.uint64), oc)
        # variable-length batch, sorted positions
        var = [bytes(src[s0:s0 + int(rng.integers(1, min(n - s0, 30) + 1))]) for s0 in rng.integers(0, n, size=64)]
        offs_v, pos_v = gpu.locate_batch(var, sorted_=True)
        for i, q in enumerate(var):
            assert np.array_equal(pos_v[int(offs_v[i]):int(offs_v[i + 1])].astype(np.uint64), np.sort(ora.locate(q))), (case, q)
        gpu.close()


def test_concurrent_callers(oracle, fm):
    """`FmIndex` is Send + Sync in the reference (immutable &self queries): concurrent batch calls on ONE handle from
    several host threads must each get their own stream + scratch arena and the right answers."""
    import threading
    po = oracle
    rng = np.random.default_rng(5)
    n = 300_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    t = po.IndexType(32, 3, 64, True)
    blob = po.build_blob(t, text, sc, table, 3, 2)
    ora = po.OracleFmIndex.load(blob, t)
    gpu = fm.FmIndex.load(blob, fm.IndexType(32, 3, 64, True))
    jobs = []
    for j in range(8):
        ln = 5 + 3 * j
        starts = rng.integers(0, n - ln, size=40_000)
        pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
        pats[::9, 1] = ord("C")
        jobs.append((pats, ora.locate_batch(pats, threads=4)))
    errors = []

    def worker(j):
        try:
            pats, (oc, oo, op_, _) = jobs[j]
            for _ in range(6):
                assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc)
                offs, pos = gpu.locate_batch(pats)
                assert np.array_equal(offs, oo) and np.array_equal(pos, op_)
        except Exception as e:  # noqa: BLE001
            errors.append((j, repr(e)))

    threads = [threading.Thread(target=worker, args=(j,)) for j in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_packed_patterns_and_u32_offsets(oracle, fm):
    """The byte-saving variants of the host entry points (include/svfm.h): 2- and 5-bit packed fixed-length patterns
    and u32 CSR offsets give exactly the results of the plain calls, across chunk sizes."""
    from sview_fmindex_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(606)
    n = 300_000
    try:
        for symbols, alphabet, bits, ln, with_wildcard in (([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"], b"ACGT", 2, 20, False),
                                                           ([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"], b"ACGTN", 3, 7, False),
                                                           ([bytes([c]) for c in b"ACDEFGHIKLMNPQRSTVWY"], b"ACDEFGHIKLMNPQRSTVWY", 5, 12, True)):
            text = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)].copy()
            planes = 3 if len(symbols) <= 7 else 5
            ora, gpu, table, sc = _pair(oracle, fm, bytes(text), symbols, 32, planes, 64, 3, 2, with_wildcard=with_wildcard)
            m = 60_000
            starts = rng.integers(0, n - ln, size=m)
            pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
            pats[::6, ln // 2] = alphabet[0]
            oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
            packed = gpu.pack_patterns(pats, table, bits)
            assert packed.shape == (m, (ln * bits + 7) // 8)
            for chunk in (0, 9000):
                assert L.svfm_set_tuning(_ffi.SVFM_TUNE_CHUNK, chunk) == 0
                assert np.array_equal(gpu.count_batch_packed(packed, ln, bits).astype(np.uint64), oc)
                offs, pos = gpu.locate_batch_packed(packed, ln, bits)
                assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64))
                offs32, pos32 = gpu.locate_batch_packed(packed, ln, bits, offs32=True)
                assert offs32.dtype == np.uint32 and np.array_equal(offs32.astype(np.uint64), oo) and np.array_equal(pos32, pos)
                o3, p3 = gpu.locate_batch(pats, offs32=True)
                assert o3.dtype == np.uint32 and np.array_equal(o3.astype(np.uint64), oo) and np.array_equal(p3, pos)
            # a symbol index that does not fit `bits` is refused by the packer; one >= symbol_count by the search
            with pytest.raises(fm.SvfmError):
                gpu.pack_patterns(np.full((4, ln), 255, dtype=np.uint8), None, bits)
            gpu.close()
    finally:
        L.svfm_set_tuning(_ffi.SVFM_TUNE_CHUNK, _ffi.SVFM_TUNE_AUTO)


def test_long_patterns_text_verification(oracle, fm):
    """Long patterns on texts with repeats: the search switches to text verification as soon as a few candidate rows are
    left (search_kernels.cuh), which must give the reference's answers -- unique hits, absent patterns (mismatch far from the
    seed), patterns that run into the start of the text, true repeats of the whole pattern (SA-row order!), and reads over
    repeated regions with one mutated copy."""
    rng = np.random.default_rng(1505)
    unit = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=700)]
    body = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=60_000)]
    mutated = unit.copy()
    mutated[350] = ord("A") if mutated[350] != ord("A") else ord("C")
    text = np.concatenate([unit, body[:20_000], unit, body[20_000:40_000], mutated, body[40_000:], unit[:300]])
    n = len(text)
    for (p, nn, v, k, r, syms) in ((32, 2, 64, 3, 2, [b"A", b"C", b"G", b"T"]), (32, 3, 64, 3, 16, [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]),
                                   (64, 3, 128, 2, 5, [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]), (64, 4, 32, 4, 1, [b"A", b"C", b"G", b"T"])):
        ora, gpu, _, _ = _pair(oracle, fm, bytes(text), syms, p, nn, v, k, r)
        for ln in (33, 64, 150, 400):
            m = 600
            starts = rng.integers(0, n - ln, size=m)
            starts[:40] = rng.integers(0, 700 - min(ln, 600), size=40)             # inside the first copy of the repeat
            starts[40:60] = 0                                                       # at the very start of the text
            pats = text[starts[:, None] + np.arange(ln)[None, :]].copy()
            pats[60:120, 0] = ord("G")                                              # mismatch at the far end (consumed last)
            pats[120:160, ln // 2] = ord("T")
            pats[160:170] = text[:ln]
            pats[160:170, 0] = ord("C") if text[0] != ord("C") else ord("A")        # would need a symbol before the text
            oc, oo, op_, _ = ora.locate_batch(pats, threads=4)
            assert int(oc.max()) >= 2 and int(oc.min()) == 0
            assert np.array_equal(gpu.count_batch(pats).astype(np.uint64), oc), (p, nn, v, ln)
            offs, pos = gpu.locate_batch(pats)
            assert np.array_equal(offs, oo) and np.array_equal(pos.astype(np.uint64), op_.astype(np.uint64)), (p, nn, v, ln)
            var = [bytes(q[: 20 + (i * 7) % (ln - 19)]) for i, q in enumerate(pats[:200])]
            o2, p2 = gpu.locate_batch(var)
            for i in (0, 3, 41, 55, 61, 130, 165, 199):
                assert np.array_equal(p2[int(o2[i]):int(o2[i + 1])].astype(np.uint64), ora.locate(var[i])), (ln, i)
            for q in (pats[0], pats[45], pats[70], pats[165]):                      # single-pattern calls (small-batch kernel)
                assert gpu.count(bytes(q)) == ora.count(bytes(q))
                assert np.array_equal(gpu.locate(bytes(q)).astype(np.uint64), ora.locate(bytes(q)))
        gpu.close()
