"""CPU tests that pin the ORACLE (oracle/) before anything is compared against it.

They follow the reference's own test strategy:
  * README goldens                       sview-fmindex/src/tests/readme/mod.rs:15,33,37,44
  * accuracy vs an independent matcher   sview-fmindex/src/tests/get_accurate_result/mod.rs:61-142
  * config invariance                    sview-fmindex/src/tests/config_invariance/mod.rs:51-143
  * slice == rev-iter == PassThrough     sview-fmindex/src/tests/text_encoders_consistency/mod.rs:111-178
plus the edge cases of SURVEY.md Appendix B and the LoadError / BuildError paths.
"""
import json
import os

import numpy as np
import pytest

from helpers import (ALL_TYPES, brute_force_locate, gen_rand_chr_list, gen_rand_pattern, gen_rand_text)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "readme_example.json")


def _index(po, text, chr_list, p, n, v, k, r, passthrough=False):
    """Build + load like assert_accurate_fm_index (get_accurate_result/mod.rs:24-58)."""
    groups = [bytes([c]) for c in chr_list]
    table, sc = po.encoding_table(groups)
    if passthrough:
        t = po.IndexType(p, n, v, False)
        enc = table[np.frombuffer(text, dtype=np.uint8)]
        blob = po.build_blob(t, enc, sc, None, k, r)
    else:
        t = po.IndexType(p, n, v, True)
        blob = po.build_blob(t, text, sc, table, k, r)
    return po.OracleFmIndex.load(blob, t), table, sc


def test_readme_golden(oracle):
    po = oracle
    g = json.load(open(GOLDEN))
    table, sc = po.encoding_table([s.encode() for s in g["symbols"]])
    assert sc == 4
    t = po.IndexType(**g["type"])
    blob = po.build_blob(t, g["text"].encode(), sc, table, g["kmer_size"], g["sampling_ratio"])
    assert blob.size == g["blob_size"]
    ix = po.OracleFmIndex.load(blob, t)
    assert ix.layout.header_size == g["header_size"]
    for q in g["queries"]:
        pat = q["pattern"].encode()
        assert ix.count(pat) == q["count"]
        assert sorted(int(x) for x in ix.locate(pat)) == q["locate_sorted"]
        assert ix.count_rev_iter(pat[::-1]) == q["count"]
        assert sorted(int(x) for x in ix.locate_rev_iter(pat[::-1])) == q["locate_sorted"]


def test_readme_blob_bytes(oracle):
    """Layout known-answer (independent model, see the fixture's provenance)."""
    po = oracle
    g = json.load(open(GOLDEN))
    table, sc = po.encoding_table([s.encode() for s in g["symbols"]])
    t = po.IndexType(**g["type"])
    blob = po.build_blob(t, g["text"].encode(), sc, table)
    L = po.blob_layout(t, len(g["text"]), sc)
    for name, off in g["offsets"].items():
        assert getattr(L, "off_" + name) == off, name
    assert blob[:8].tobytes() == b"FI00\0\0\0\0"
    assert np.array_equal(blob[8:264], table)
    assert list(np.frombuffer(blob[264:280].tobytes(), dtype=np.uint32)) + \
        [int(np.frombuffer(blob[280:288].tobytes(), dtype=np.uint64)[0])] == g["count_array_header"]
    assert blob[g["header_size"]:].tobytes().hex() == g["body_hex"]
    sa = np.frombuffer(blob[384:384 + 31 * 4].tobytes(), dtype=np.uint32)
    assert list(sa) == g["suffix_array"]
    planes = np.frombuffer(blob[536:552].tobytes(), dtype=np.uint64)
    assert [hex(int(x)) for x in planes] == g["planes_u64"]


def test_suffix_array_vs_naive(oracle):
    po = oracle
    rng = np.random.default_rng(7)
    cases = [np.array([3], dtype=np.uint8), np.array([1, 1, 1, 1, 1], dtype=np.uint8),
             np.array([2, 1] * 40, dtype=np.uint8), np.array([1, 2, 3] * 33, dtype=np.uint8)]
    for _ in range(200):
        n = int(rng.integers(1, 300))
        K = int(rng.integers(1, 6))
        cases.append(rng.integers(1, K + 1, size=n).astype(np.uint8))
    for s in cases:
        s1 = np.concatenate([s, [0]]).astype(np.uint8)
        sa = po.suffix_array(s1, int(s.max()))
        b = s1.tobytes()
        assert list(sa) == sorted(range(len(b)), key=lambda i: b[i:])


@pytest.mark.parametrize("chr_count", [3, 4, 5, 7, 8, 9, 15, 16, 17, 33, 63, 64])
def test_results_are_accurate(oracle, chr_count):
    """get_accurate_result/mod.rs:61-142: ltks=3, sasr=2, every (P, Block, Vector) that can hold the alphabet."""
    po = oracle
    rng = np.random.default_rng(1000 + chr_count)
    for _ in range(2):
        chr_list = gen_rand_chr_list(rng, chr_count)
        text = gen_rand_text(rng, chr_list, 100, 300)
        patterns = [gen_rand_pattern(rng, text, 1, 10) for _ in range(100)]
        table, sc = po.encoding_table([bytes([c]) for c in chr_list])
        enc_text = table[np.frombuffer(text, dtype=np.uint8)]
        answers = [brute_force_locate(enc_text, table[np.frombuffer(p, dtype=np.uint8)]) for p in patterns]
        tested = 0
        for (p, n, v) in ALL_TYPES:
            if (1 << n) < chr_count:
                with pytest.raises(po.OracleError) as e:
                    _index(po, text, chr_list, p, n, v, 3, 2)
                assert e.value.code == po.ORA_ERR_SYMBOL_COUNT_OVER
                assert e.value.detail == (1 << n, chr_count)
                continue
            ix, _, _ = _index(po, text, chr_list, p, n, v, 3, 2)
            for pat, ans in zip(patterns, answers):
                got = np.sort(ix.locate(pat))
                assert np.array_equal(got, ans), (p, n, v, pat)
                assert ix.count(pat) == len(ans)
            tested += 1
        assert tested > 0


def test_config_invariance(oracle):
    """config_invariance/mod.rs:51-143: kLTS in {None,2,3,4} x SA in {Uncompressed,2,3,4} x every type."""
    po = oracle
    rng = np.random.default_rng(42)
    chr_list = b"ACGT"
    text = gen_rand_text(rng, chr_list, 1000, 1000)
    patterns = [gen_rand_pattern(rng, text, 10, 10)] + [gen_rand_pattern(rng, text, 1, 6) for _ in range(5)]
    base, table, _ = _index(po, text, chr_list, 32, 4, 32, 1, 1)
    answers = [np.sort(base.locate(p)) for p in patterns]
    enc_text = table[np.frombuffer(text, dtype=np.uint8)]
    for pat, ans in zip(patterns, answers):
        assert np.array_equal(ans, brute_force_locate(enc_text, table[np.frombuffer(pat, dtype=np.uint8)]))
    for k in (1, 2, 3, 4):
        for r in (1, 2, 3, 4):
            for (p, n, v) in ALL_TYPES:
                ix, _, _ = _index(po, text, chr_list, p, n, v, k, r)
                for pat, ans in zip(patterns, answers):
                    assert np.array_equal(np.sort(ix.locate(pat)), ans), (k, r, p, n, v, pat)


def test_text_encoders_are_consistent(oracle):
    """text_encoders_consistency/mod.rs:86-106: slice == rev iter, EncodingTable == PassThrough on encoded text."""
    po = oracle
    rng = np.random.default_rng(3)
    for chr_count in (2, 4, 5, 21):
        chr_list = gen_rand_chr_list(rng, chr_count)
        text = gen_rand_text(rng, chr_list, 200, 400)
        for (p, n, v) in [(32, 5, 64), (64, 6, 128), (32, 6, 32)] if chr_count > 8 else [(32, 3, 64), (64, 4, 32), (32, 5, 128)]:
            ixt, table, sc = _index(po, text, chr_list, p, n, v, 3, 2)
            ixp, _, _ = _index(po, text, chr_list, p, n, v, 3, 2, passthrough=True)
            for _ in range(60):
                pat = gen_rand_pattern(rng, text, 1, 12)
                enc = bytes(table[np.frombuffer(pat, dtype=np.uint8)])
                c = ixt.count(pat)
                assert c == ixt.count_rev_iter(pat[::-1])
                assert c == ixp.count(enc) == ixp.count_rev_iter(enc[::-1])
                a = np.sort(ixt.locate(pat))
                assert np.array_equal(a, np.sort(ixt.locate_rev_iter(pat[::-1])))
                assert np.array_equal(a, np.sort(ixp.locate(enc)))
                assert np.array_equal(a, np.sort(ixp.locate_rev_iter(enc[::-1])))
                # identical SA interval, and identical (unsorted) SA-row order
                assert ixt.pos_range(pat) == ixt.pos_range(pat[::-1], True) == ixp.pos_range(enc)
                assert np.array_equal(ixt.locate(pat), ixp.locate(enc))


@pytest.mark.parametrize("vec_bits", [32, 64, 128])
def test_edge_cases(oracle, vec_bits):
    """SURVEY.md Appendix B."""
    po = oracle
    rng = np.random.default_rng(11 + vec_bits)
    chr_list = b"ACGTN"
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    assert sc == 5
    # n % BLOCK_LEN == 0 (extra all-zero block), n = BLOCK_LEN - 1, BLOCK_LEN + 1, tiny texts
    for n in (vec_bits * 4, vec_bits, vec_bits - 1, vec_bits + 1, 1, 2, 3):
        text = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)])
        for k, r in ((1, 1), (3, 2), (4, 5)):
            t = po.IndexType(32, 3, vec_bits, True)
            blob = po.build_blob(t, text, sc, table, k, r)
            assert blob.size == po.blob_layout(t, n, sc, k, r).total_size
            ix = po.OracleFmIndex.load(blob, t)
            assert ix.text_len == n
            enc_text = table[np.frombuffer(text, dtype=np.uint8)]
            pats = {text[i:i + m] for i in range(0, n, max(1, n // 7)) for m in (1, 2, 3, 4, 9) if i + m <= n}
            pats |= {b"A", b"AC", b"ACG", b"TTTTTTTT", b"N", b"NN", b"xyz", b"GATTACA"}
            for pat in pats:
                ans = brute_force_locate(enc_text, table[np.frombuffer(pat, dtype=np.uint8)])
                assert ix.count(pat) == len(ans), (n, k, r, pat)
                assert np.array_equal(np.sort(ix.locate(pat)), ans), (n, k, r, pat)
                assert ix.count_rev_iter(pat[::-1]) == len(ans)
    # homopolymer (deep recursion in the suffix sorter, every row in one symbol class)
    text = b"A" * 777
    t = po.IndexType(64, 2, vec_bits, True)
    tb4, sc4 = po.encoding_table([b"A", b"C", b"G", b"T"])
    ix = po.OracleFmIndex.load(po.build_blob(t, text, sc4, tb4, 2, 3), t)
    assert ix.count(b"AAAA") == 774
    assert np.array_equal(np.sort(ix.locate(b"AAAA")), np.arange(774, dtype=np.uint64))
    assert ix.count(b"C") == 0 and ix.locate(b"AC").size == 0
    # empty pattern: the reference panics (count_array.rs:211); the restatement reports it
    with pytest.raises(po.OracleError) as e:
        ix.count(b"")
    assert e.value.code == po.ORA_ERR_EMPTY_PATTERN


def test_load_errors(oracle):
    """load_from_blob.rs:28-58."""
    po = oracle
    table, sc = po.encoding_table([b"A", b"C", b"G", b"T"])
    t = po.IndexType(32, 2, 64, True)
    blob = po.build_blob(t, b"ACGTACGTTTGACCA", sc, table, 2, 2)
    bad = blob.copy()
    bad[0] = ord("X")
    with pytest.raises(po.OracleError) as e:
        po.OracleFmIndex.load(bad, t)
    assert e.value.code == po.ORA_ERR_INVALID_FORMAT
    bad = blob.copy()
    bad[2] = ord("1")  # unsupported version
    with pytest.raises(po.OracleError) as e:
        po.OracleFmIndex.load(bad, t)
    assert e.value.code == po.ORA_ERR_INVALID_FORMAT
    longer = po.aligned_empty(blob.size + 8)
    longer[:blob.size] = blob
    with pytest.raises(po.OracleError) as e:
        po.OracleFmIndex.load(longer, t)
    assert e.value.code == po.ORA_ERR_BLOB_SIZE and e.value.detail == (blob.size, blob.size + 8)
    # wrong type triple: sizes no longer add up
    with pytest.raises(po.OracleError) as e:
        po.OracleFmIndex.load(blob, po.IndexType(64, 2, 64, True))
    assert e.value.code == po.ORA_ERR_BLOB_SIZE
    with pytest.raises(po.OracleError):
        po.OracleFmIndex.load(blob[:100], t)


def test_blob_sizes_of_bench_configs(oracle):
    """Sizes computed in SURVEY.md section 8b (builder/mod.rs:165-181)."""
    po = oracle
    assert po.blob_layout(po.IndexType(32, 3, 64, True), 10**9, 5, 3, 2).total_size == 2_687_501_296
    assert po.blob_layout(po.IndexType(32, 2, 64, True), 10**9, 4, 3, 2).total_size == 2_500_000_920
    assert po.blob_layout(po.IndexType(32, 2, 64, True), 10**9, 4, 3, 16).total_size == 750_000_920
    assert po.blob_layout(po.IndexType(32, 5, 64, True), 5 * 10**8, 21, 3, 2).total_size == 1_968_793_168
    assert po.blob_layout(po.IndexType(64, 3, 64, True), 31 * 10**8, 5, 3, 2).total_size == 15_500_002_200
    assert po.lib().ora_kmer_size_for_max_memory(32, 1, 0) == 1  # lookup_table_config.rs:73-75
    assert po.lib().ora_kmer_size_for_max_memory(32, 4, 4 * 5**6) == 6


def test_batch_drivers_match_single_calls(oracle):
    po = oracle
    rng = np.random.default_rng(5)
    text = gen_rand_text(rng, b"ACGT", 5000, 5000)
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    t = po.IndexType(32, 3, 64, True)
    ix = po.OracleFmIndex.load(po.build_blob(t, text, sc, table, 3, 2), t)
    tarr = np.frombuffer(text, dtype=np.uint8)
    starts = rng.integers(0, len(text) - 6, size=500)
    pats = np.stack([tarr[s:s + 6] for s in starts])
    for threads in (1, 3):
        counts = ix.count_batch(pats, threads)
        c2, offs, pos, ck = ix.locate_batch(pats, threads)
        c3, _, _, ck3 = ix.locate_batch(pats, threads, want_positions=False)
        assert np.array_equal(counts, c3) and ck == ck3
        acc = 0
        for i in range(len(pats)):
            loc = ix.locate(bytes(pats[i]))
            assert counts[i] == len(loc)
            assert np.array_equal(pos[int(offs[i]):int(offs[i + 1])].astype(np.uint64), loc)
            acc = (acc + sum((int(x) + 1) * (2 * i + 1) for x in loc)) % 2**64
        assert acc == ck
