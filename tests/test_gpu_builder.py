"""GPU index construction: the blob built on the device must equal, byte for byte,
the blob the oracle's restatement of FmIndexBuilder::build produces (builder/mod.rs:187-264)."""
import ctypes as C

import numpy as np
import pytest

from helpers import ALL_TYPES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fm():
    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi
    _ffi.lib()
    return fm


def _gpu_blob(fm, text, ft, S, enc, k, r):
    b = fm.FmIndexBuilder(len(text), S, enc, ft)
    b.kmer_size, b.sampling_ratio = k, r
    blob = fm.aligned_empty(b.blob_size())
    b.build(text, blob)
    return blob


def _first_diff(a, b):
    d = np.nonzero(a != b)[0]
    return int(d[0]) if d.size else -1


def test_small_texts_every_type(oracle, fm):
    po = oracle
    rng = np.random.default_rng(77)
    for (p, n, v) in ALL_TYPES:
        S = int(rng.integers(1, (1 << n) + 1))
        syms = [bytes([65 + i]) for i in range(S)]
        enc = fm.EncodingTable.from_symbols(syms)
        table, sc = po.encoding_table(syms)
        assert sc == S
        for tl in (1, 2, v - 1, v, v + 1, 5 * v, int(rng.integers(200, 3000))):
            text = (65 + rng.integers(0, S, size=tl)).astype(np.uint8)
            k, r = int(rng.integers(1, 4)), int(rng.integers(1, 6))
            exp = po.build_blob(po.IndexType(p, n, v, True), text, S, table, k, r)
            got = _gpu_blob(fm, text, fm.IndexType(p, n, v, True), S, enc, k, r)
            assert got.size == exp.size
            assert np.array_equal(got, exp), (p, n, v, S, tl, k, r, _first_diff(got, exp))


def test_repetitive_texts_need_prefix_doubling(oracle, fm):
    """Long repeats: ties after the packed-prefix sort are resolved by the prefix-doubling rounds."""
    po = oracle
    rng = np.random.default_rng(5)
    syms = [b"A", b"C", b"G", b"T"]
    enc = fm.EncodingTable.from_symbols(syms)
    table, S = po.encoding_table(syms)
    unit = np.frombuffer(b"ACGTTGCA" * 40, dtype=np.uint8)
    texts = [np.full(5000, 65, dtype=np.uint8),
             np.tile(np.frombuffer(b"AC", dtype=np.uint8), 3000),
             np.tile(unit, 30),
             np.concatenate([np.tile(unit, 7), np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 999)], np.tile(unit, 9)])]
    for text in texts:
        for (p, n, v, k, r) in ((32, 2, 64, 2, 2), (64, 3, 128, 1, 3)):
            exp = po.build_blob(po.IndexType(p, n, v, True), text, S, table, k, r)
            got = _gpu_blob(fm, text, fm.IndexType(p, n, v, True), S, enc, k, r)
            assert np.array_equal(got, exp), (len(text), p, n, v, _first_diff(got, exp))


def test_bench_shaped_index_and_passthrough(oracle, fm):
    """cfg1 shape at 3 Mbp (u32, Block3<u64>, S=5, r=2, k=3) and a protein-shaped index (S=21 with wildcard,
    Block5<u64>), plus PassThrough on pre-encoded text."""
    po = oracle
    from sview_fmindex_b200 import synth
    text = synth.synth_text(3_000_000, 42, synth.NUCLEOTIDES)
    syms = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]
    table, S = po.encoding_table(syms)
    exp = po.build_blob(po.IndexType(32, 3, 64, True), text, S, table, 3, 2)
    got = _gpu_blob(fm, text, fm.IndexType(32, 3, 64, True), S, fm.EncodingTable.from_symbols(syms), 3, 2)
    assert np.array_equal(got, exp), _first_diff(got, exp)

    prot = synth.synth_text(700_000, 42, synth.AMINO_ACIDS, rare=1000, rare_byte=ord("X"))
    psyms = [bytes([c]) for c in synth.AMINO_ACIDS]
    ptable, pS = po.encoding_table(psyms, with_wildcard=True)
    assert pS == 21
    for (n, v) in ((5, 64), (6, 64)):
        exp = po.build_blob(po.IndexType(32, n, v, True), prot, pS, ptable, 3, 2)
        got = _gpu_blob(fm, prot, fm.IndexType(32, n, v, True), pS, fm.EncodingTable.from_symbols_with_wildcard(psyms), 3, 2)
        assert np.array_equal(got, exp), (n, v, _first_diff(got, exp))
    enc_text = ptable[prot]
    exp = po.build_blob(po.IndexType(64, 5, 32, False), enc_text, pS, None, 2, 4)
    got = _gpu_blob(fm, enc_text, fm.IndexType(64, 5, 32, False), pS, None, 2, 4)
    assert np.array_equal(got, exp), _first_diff(got, exp)
    # the built blob loads and answers like the oracle
    ix = fm.FmIndex.load(got, fm.IndexType(64, 5, 32, False))
    ora = po.OracleFmIndex.load(exp, po.IndexType(64, 5, 32, False))
    pats, _ = synth.synth_patterns(enc_text, 2000, 5, 7)
    oc, oo, op_, _ = ora.locate_batch(pats, threads=2)
    offs, pos = ix.locate_batch(pats)
    assert np.array_equal(offs, oo) and np.array_equal(pos, op_)


def test_builder_errors(fm):
    enc = fm.EncodingTable.from_symbols([b"A", b"C", b"G", b"T"])
    b = fm.FmIndexBuilder(100, 4, enc, fm.IndexType(32, 2, 64, True))
    text = np.full(100, 65, dtype=np.uint8)
    with pytest.raises(fm.BuildError) as e:
        b.build(text, fm.aligned_empty(b.blob_size() + 8))
    assert e.value.code == 12 and e.value.detail == (b.blob_size(), b.blob_size() + 8)
    with pytest.raises(fm.BuildError) as e:
        b.build(text[:50], fm.aligned_empty(b.blob_size()))
    assert e.value.code == 11
    # PassThrough text with a byte >= symbol_count
    bp = fm.FmIndexBuilder(100, 4, None, fm.IndexType(32, 2, 64, False))
    with pytest.raises(fm.SvfmError) as e:
        bp.build(text, fm.aligned_empty(bp.blob_size()))
    assert e.value.code == 24


def test_synth_device_matches_numpy_twin(fm):
    import torch
    from sview_fmindex_b200 import _ffi, synth
    L = _ffi.lib()
    n = 100_003
    d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
    alpha = np.frombuffer(synth.AMINO_ACIDS, dtype=np.uint8)
    assert L.svfm_bench_synth_text(d_text.data_ptr(), n, 42, alpha.ctypes.data, len(alpha), 1000, ord("X"), None) == 0
    assert np.array_equal(d_text.cpu().numpy(), synth.synth_text(n, 42, synth.AMINO_ACIDS, 1000, ord("X")))
    m, ln = 5000, 12
    d_pats = torch.empty(m * ln, dtype=torch.uint8, device="cuda")
    d_starts = torch.empty(m, dtype=torch.int64, device="cuda")
    assert L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pats.data_ptr(), d_starts.data_ptr(), m, ln, 9, None) == 0
    torch.cuda.synchronize()
    pats, starts = synth.synth_patterns(d_text.cpu().numpy(), m, ln, 9)
    assert np.array_equal(d_pats.cpu().numpy().reshape(m, ln), pats)
    assert np.array_equal(d_starts.cpu().numpy().astype(np.uint64), starts)
