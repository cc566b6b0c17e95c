"""Blob byte parity with the REAL reference builder, for whenever somebody with cargo has run
bindings/rust/sview-fmindex-b200/examples/dump_fixtures.rs and committed its output under tests/golden/rust_blobs/.
Every fixture = a blob the Rust crate built + the answers its `count` / `locate` gave.  Checked here:
  * the oracle builder (CPU restatement) reproduces the blob byte for byte,            [CPU]
  * the oracle's count / locate on the Rust-built blob give the Rust answers,            [CPU]
  * the GPU builder reproduces the blob, and the GPU search on it gives the Rust answers [GPU]
Skipped while the directory holds no fixture (this image has no Rust toolchain)."""
import glob
import json
import os

import numpy as np
import pytest

DIR = os.environ.get("SVFM_RUST_FIXTURES") or os.path.join(os.path.dirname(__file__), "golden", "rust_blobs")
FIXTURES = sorted(glob.glob(os.path.join(DIR, "*.json")))
needs_fixtures = pytest.mark.skipif(not FIXTURES, reason="no Rust-built fixture under tests/golden/rust_blobs (see its README)")


def _load(path):
    from sview_fmindex_b200 import synth
    meta = json.load(open(path))
    blob = np.fromfile(os.path.join(DIR, meta["blob_file"]), dtype=np.uint8)
    assert blob.size == meta["blob_len"]
    text = synth.synth_text(meta["text_len"], meta["text_seed"], meta["alphabet"].encode())
    return meta, blob, text


@needs_fixtures
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_oracle_reproduces_the_rust_blob(oracle, path):
    po = oracle
    meta, blob, text = _load(path)
    t = po.IndexType(meta["pos_bits"], meta["planes"], meta["vec_bits"], True)
    table, sc = po.encoding_table([s.encode() for s in meta["symbols"]], meta["wildcard"])
    mine = po.build_blob(t, text, sc, table, meta["kmer_size"], meta["sampling_ratio"])
    assert mine.size == blob.size
    diff = np.flatnonzero(mine != blob)
    assert diff.size == 0, f"first differing byte at offset {int(diff[0])}"
    aligned = po.aligned_empty(blob.size)
    aligned[:] = blob
    ora = po.OracleFmIndex.load(aligned, t)
    for q in meta["queries"]:
        pat = bytes.fromhex(q["pattern_hex"])
        assert ora.count(pat) == q["count"]
        assert [int(x) for x in ora.locate(pat)] == q["locate"]      # SA-row order, as the Rust crate returned it


@needs_fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_gpu_matches_the_rust_blob(path):
    import sview_fmindex_b200 as fm
    meta, blob, text = _load(path)
    it = fm.IndexType(meta["pos_bits"], meta["planes"], meta["vec_bits"], True)
    syms = [s.encode() for s in meta["symbols"]]
    enc = fm.EncodingTable.from_symbols_with_wildcard(syms) if meta["wildcard"] else fm.EncodingTable.from_symbols(syms)
    b = fm.FmIndexBuilder(text.size, enc.symbol_count(), enc, it)
    b.kmer_size, b.sampling_ratio = meta["kmer_size"], meta["sampling_ratio"]
    mine = fm.aligned_empty(b.blob_size())
    b.build(text, mine)
    assert np.array_equal(mine, blob)
    ix = fm.FmIndex.load(blob, it)
    pats = [bytes.fromhex(q["pattern_hex"]) for q in meta["queries"]]
    counts = ix.count_batch(pats)
    offs, pos = ix.locate_batch(pats)
    for i, q in enumerate(meta["queries"]):
        assert int(counts[i]) == q["count"]
        assert [int(x) for x in pos[int(offs[i]):int(offs[i + 1])]] == q["locate"]
