"""Bench-CLI-compatible front end (sview_fmindex_b200/bench_cli.py; reference bench/src/{main,build,locate}): file
formats on the CPU, and the build -> locate round trip on the GPU against the oracle."""
import os

import numpy as np
import pytest


def test_pattern_reader_and_result_format(tmp_path):
    from sview_fmindex_b200 import bench_cli
    p = tmp_path / "pattern.txt"
    p.write_bytes(b"ACGT\nTTGA\r\nA\nGATTACA")          # last line without terminator, one CRLF
    data, offs = bench_cli.read_patterns(str(p))
    assert list(offs) == [0, 4, 8, 9, 16]
    assert bytes(data) == b"ACGTTTGAAGATTACA"
    (tmp_path / "empty.txt").write_bytes(b"")
    data, offs = bench_cli.read_patterns(str(tmp_path / "empty.txt"))
    assert data.size == 0 and list(offs) == [0]
    # bench/src/locate/mod.rs:115-123: positions joined by ',', one line per pattern, empty line for no occurrence
    out = bench_cli.format_results(np.array([0, 2, 2, 3], dtype=np.uint64), np.array([5, 18, 7], dtype=np.uint32))
    assert out == b"5,18\n\n7\n"
    assert bench_cli.format_results(np.array([0], dtype=np.uint64), np.zeros(0, dtype=np.uint32)) == b""
    assert bench_cli.blob_stem(True) == "sview-memory-block2" and bench_cli.blob_stem(False) == "sview-memory-block3"


@pytest.mark.gpu
@pytest.mark.parametrize("t_wild", [False, True])
def test_build_locate_round_trip(oracle, tmp_path, t_wild):
    """generate -> build -> locate through the CLI; the blob file must be the oracle builder's bytes and every result
    line the oracle's `locate` in SA-row order (what the reference binary would have written)."""
    from sview_fmindex_b200 import bench_cli
    po = oracle
    d = str(tmp_path)
    assert bench_cli.main(["generate-text", "-d", d, "-t", "200000", "-s", "42"]) == 0
    assert bench_cli.main(["generate-pattern", "-d", d, "-p", "9", "-n", "3000", "-s", "42"]) == 0
    # add patterns that do not occur, lower case (symbols Aa..), N, and variable lengths
    with open(os.path.join(d, "pattern.txt"), "ab") as f:
        f.write(b"acgtacgtac\nNNNN\nGATTACAGATTACAGATTACA\nT\nCG\n")
    args = ["-d", d] + (["-t"] if t_wild else [])
    assert bench_cli.main(["build", "-s", "2", "-k", "3"] + args) == 0
    assert bench_cli.main(["locate"] + args) == 0
    stem = bench_cli.blob_stem(t_wild)
    text = np.fromfile(os.path.join(d, "text.txt"), dtype=np.uint8)
    table, sc = po.encoding_table(bench_cli.SYMBOLS_ACGT if t_wild else bench_cli.SYMBOLS_ACGTN)
    t = po.IndexType(32, 2 if t_wild else 3, 64, True)
    blob = po.build_blob(t, text, sc, table, 3, 2)
    assert np.array_equal(np.fromfile(os.path.join(d, stem + ".blob"), dtype=np.uint8), blob)
    ora = po.OracleFmIndex.load(blob, t)
    pats = open(os.path.join(d, "pattern.txt"), "rb").read().split(b"\n")[:-1]
    lines = open(os.path.join(d, stem + "-results.txt"), "rb").read().split(b"\n")[:-1]
    assert len(lines) == len(pats) == 3005
    for pat, line in zip(pats, lines):
        assert line == b",".join(str(int(x)).encode() for x in ora.locate(pat)), pat
    # the mmap variant (bench/src/locate/sview_mmap.rs) loads the same blob out of the page cache: identical result file
    first = open(os.path.join(d, stem + "-results.txt"), "rb").read()
    os.remove(os.path.join(d, stem + "-results.txt"))
    os.environ["MMAP_ADVICE_SEQUENTIAL"] = "1"
    try:
        assert bench_cli.main(["locate", "-a", "sview-mmap"] + args) == 0
    finally:
        del os.environ["MMAP_ADVICE_SEQUENTIAL"]
    assert open(os.path.join(d, stem + "-results.txt"), "rb").read() == first


def test_mmap_blob_is_a_read_only_mapping(tmp_path):
    from sview_fmindex_b200 import bench_cli
    p = tmp_path / "x.blob"
    p.write_bytes(bytes(range(256)) * 64)
    os.environ["MMAP_ADVICE_RANDOM"] = "1"
    try:
        m = bench_cli.map_blob(str(p))
    finally:
        del os.environ["MMAP_ADVICE_RANDOM"]
    assert isinstance(m, np.memmap) and m.size == 256 * 64 and int(m[257]) == 1 and not m.flags.writeable
