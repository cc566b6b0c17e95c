"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/svfm.h declares,
the host-only parts (blob validation, blob_size, error mapping) behave like the reference, and the
product fails loudly without a CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fm():
    import __graft_entry__ as ge
    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        ge.build()
    return fm


def test_library_exports_every_declared_symbol(fm):
    from sview_fmindex_b200 import _ffi
    for header, table, path in (("svfm.h", _ffi.EXPORTS, _ffi.LIB_PATH), ("svfm_bench.h", _ffi.BENCH_EXPORTS, _ffi.BENCH_LIB_PATH)):
        lib = C.CDLL(path)
        hdr = open(os.path.join(ROOT, "include", header)).read()
        declared = set(re.findall(r"\b(svfm_[a-z0-9_]+)\s*\(", hdr))
        assert declared
        for name in sorted(declared):
            assert hasattr(lib, name), f"{os.path.basename(path)} does not export {name}"
        bound = {e[0] for e in table}
        assert declared == bound, (header, declared - bound, bound - declared)
    # the measurement helpers are not part of the product library
    product = C.CDLL(_ffi.LIB_PATH)
    assert not any(hasattr(product, e[0]) for e in _ffi.BENCH_EXPORTS)
    assert b"sm_100a" in _ffi.lib().svfm_version()


def test_python_constants_match_the_header():
    """Every enumerator / macro of include/svfm.h that the ctypes layer mirrors has the header's value."""
    from sview_fmindex_b200 import _ffi
    hdr = open(os.path.join(ROOT, "include", "svfm.h")).read()
    consts = {k: int(v, 0) for k, v in re.findall(r"\b(SVFM_[A-Z0-9_]+)\s*=\s*(0x[0-9a-fA-F]+|\d+)", hdr)}
    consts.update({k: int(v.rstrip("ulUL"), 0) for k, v in re.findall(r"#define\s+(SVFM_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+[ulUL]*|\d+[ulUL]*)", hdr)})
    assert len(consts) > 25
    mirrored = {k: v for k, v in vars(_ffi).items() if k.startswith("SVFM_") and isinstance(v, int)}
    assert mirrored
    for k, v in mirrored.items():
        assert k in consts, f"{k} is not in include/svfm.h"
        assert consts[k] == v, (k, consts[k], v)
    for k in consts:
        if k.startswith(("SVFM_ERR_", "SVFM_TUNE_")) or k in ("SVFM_OK", "SVFM_REVERSED", "SVFM_SORTED"):
            assert k in mirrored, f"{k} is missing from sview_fmindex_b200/_ffi.py"


def test_rust_sys_source_declares_the_whole_header():
    """bindings/rust/svfm-sys is source only (no Rust toolchain in this image): at least keep its `extern "C"` block in
    step with include/svfm.h, function by function."""
    hdr = open(os.path.join(ROOT, "include", "svfm.h")).read()
    declared = set(re.findall(r"\b(svfm_[a-z0-9_]+)\s*\(", hdr))
    rs = open(os.path.join(ROOT, "bindings", "rust", "svfm-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (svfm_[a-z0-9_]+)\s*\(", rs))
    assert declared == bound, (declared - bound, bound - declared)
    for k, v in re.findall(r"\b(SVFM_TUNE_[A-Z_]+)\s*=\s*(\d+)\b", hdr):
        if k != "SVFM_TUNE_AUTO":  # a #define (u64 sentinel), not an enumerator
            assert re.search(rf"pub const {k}: c_int = {v};", rs), k


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, link or load it."""
    pkg = os.path.join(ROOT, "sview_fmindex_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dp, f), errors="replace").read()
                for needle in ("pyoracle", "fm_oracle", "libfm_oracle", "from oracle", "import oracle"):
                    assert needle not in src.replace("never imported", ""), (f, needle)


def test_blob_size_matches_reference_layout(fm, oracle):
    po = oracle
    enc = fm.EncodingTable.from_symbols([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    assert enc.symbol_count() == 5
    b = fm.FmIndexBuilder.new(10**9, 5, enc, fm.IndexType(32, 3, 64, True)) \
        .set_lookup_table_config(fm.LookupTableConfig.KmerSize(3)) \
        .set_suffix_array_config(fm.SuffixArrayConfig.Compressed(2))
    assert b.blob_size() == 2_687_501_296
    rng = np.random.default_rng(0)
    for _ in range(200):
        p = int(rng.choice([32, 64])); n = int(rng.integers(2, 7)); v = int(rng.choice([32, 64, 128]))
        e = bool(rng.integers(0, 2)); S = int(rng.integers(1, (1 << n) + 1)); k = int(rng.integers(1, 4))
        r = int(rng.integers(1, 9)); tl = int(rng.integers(1, 10**7))
        ft = fm.IndexType(p, n, v, e)
        bb = fm.FmIndexBuilder(tl, S, fm.EncodingTable(np.zeros(256, dtype=np.uint8)) if e else None, ft)
        bb.kmer_size, bb.sampling_ratio = k, r
        assert bb.blob_size() == po.blob_layout(po.IndexType(p, n, v, e), tl, S, k, r).total_size
    with pytest.raises(fm.BuildError) as ei:
        fm.FmIndexBuilder.new(100, 5, fm.EncodingTable.from_symbols([b"A"] * 5), fm.IndexType(32, 2, 64, True))
    assert ei.value.code == 10 and ei.value.detail == (4, 5)  # SymbolCountOver(max, got), builder/mod.rs:71-73
    with pytest.raises(fm.BuildError):
        fm.FmIndexBuilder.new(100, 4, enc, fm.IndexType(32, 3, 64, True)).set_lookup_table_config(fm.LookupTableConfig.KmerSize(1))
    with pytest.raises(fm.BuildError):
        fm.FmIndexBuilder.new(100, 4, enc, fm.IndexType(32, 3, 64, True)).set_suffix_array_config(fm.SuffixArrayConfig.Compressed(1))
    mm = fm.FmIndexBuilder.new(100, 4, enc, fm.IndexType(32, 3, 64, True)).set_lookup_table_config(fm.LookupTableConfig.MaxMemory(4 * 5**6))
    assert mm.kmer_size == 6


def test_encoding_table_mirror(fm, oracle):
    for syms in ([b"Aa", b"Cc", b"Gg", b"Tt"], [b"ACD", b"x"], [bytes([i]) for i in range(40, 61)]):
        for wc in (False, True):
            mine = fm.EncodingTable.from_symbols_with_wildcard(syms) if wc else fm.EncodingTable.from_symbols(syms)
            tbl, sc = oracle.encoding_table(syms, wc)
            assert np.array_equal(mine.table, tbl) and mine.symbol_count() == sc


def test_check_blob_is_the_load_error_path(fm, oracle):
    po = oracle
    table, sc = po.encoding_table([b"A", b"C", b"G", b"T"])
    for (p, n, v) in ((32, 2, 64), (64, 3, 128), (32, 6, 32)):
        blob = po.build_blob(po.IndexType(p, n, v, True), b"ACGTACGTTTGACCAGGATTACA", sc, table, 2, 3)
        ft = fm.IndexType(p, n, v, True)
        info = fm.FmIndex.check_blob(blob, ft)
        assert (info.text_len, info.symbol_count, info.kmer_size, info.sampling_ratio) == (23, 4, 2, 3)
        bad = blob.copy(); bad[3] = ord("9")
        with pytest.raises(fm.InvalidFormat):
            fm.FmIndex.check_blob(bad, ft)
        with pytest.raises(fm.MismatchedBlobSize) as e:
            fm.FmIndex.check_blob(blob[:-8 if v != 128 else -16], ft)
        assert e.value.expected == blob.size
        with pytest.raises(fm.InvalidFormat):
            fm.FmIndex.check_blob(blob[:40], ft)


def test_no_cpu_fallback(fm, oracle):
    """Without a CUDA device every compute entry point must fail loudly (SVFM_ERR_CUDA), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    po = oracle
    table, sc = po.encoding_table([b"A", b"C", b"G", b"T"])
    blob = po.build_blob(po.IndexType(32, 2, 64, True), b"ACGTACGT", sc, table)
    with pytest.raises(fm.SvfmError) as e:
        fm.FmIndex.load(blob, fm.IndexType(32, 2, 64, True))
    assert e.value.code == 30


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) on a tiny text without a GPU: one JSON line
    with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--text-len", "300000",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "count+locate patterns/s" and d["unit"] == "patterns/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_roofline_traffic_file_names_existing_kernels():
    """profiles/roofline_traffic.json (ncu DRAM bytes per kernel, read by bench.py) is regenerated by tools/roofline_traffic.py
    after kernel changes; a kernel name in it that no longer exists in the sources means the file has gone stale."""
    import json
    import sys
    doc = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    assert doc["configs"], "no configuration in profiles/roofline_traffic.json"
    src = ""
    csrc = os.path.join(ROOT, "sview_fmindex_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")):
            src += open(os.path.join(csrc, f), errors="replace").read()
    sys.path.insert(0, ROOT)
    import bench
    for name, cfg in doc["configs"].items():
        assert name in bench.CONFIGS, name
        assert cfg["patterns_per_step"] == bench.CONFIGS[name]["batch"], (name, "batch size changed: re-profile")
        phases = {k["phase"] for k in cfg["kernels"]}
        assert phases <= set(bench.PHASES) | {"other"}, phases
        for k in cfg["kernels"]:
            if k["kernel"].startswith("svfm::"):
                fn = k["kernel"].split("::")[-1]
                assert re.search(rf"\b{fn}\b", src), f"{name}: kernel {fn} is not in csrc any more -- regenerate the traffic file"


def test_kmer_table_length_overflow_is_an_invalid_config():
    """(S+1)^k is a u32 in the reference (count_array.rs:70, u32::pow): KmerSize(14) on a 4-symbol alphabet needs 5^14 =
    6.1e9 entries.  The reference panics or wraps; here the builder refuses the configuration, and MaxMemory never picks it."""
    import sview_fmindex_b200 as fm
    enc = fm.EncodingTable.from_symbols([b"A", b"C", b"G", b"T"])
    it = fm.IndexType(32, 2, 64, True)
    b = fm.FmIndexBuilder(1000, 4, enc, it)
    with pytest.raises(fm.BuildError) as e:
        b.set_lookup_table_config(fm.LookupTableConfig.KmerSize(14))
    assert e.value.code == 14
    b = fm.FmIndexBuilder(1000, 4, enc, it)
    b.set_lookup_table_config(fm.LookupTableConfig.KmerSize(13))      # 5^13 = 1.2e9 fits
    assert b.kmer_size == 13
    b.set_lookup_table_config(fm.LookupTableConfig.MaxMemory(32 << 30))  # 5^14 * 4 B = 24.4 GB would fit the budget
    assert b.kmer_size == 13
    b.set_lookup_table_config(fm.LookupTableConfig.MaxMemory(1 << 20))   # reference arithmetic below the bound
    assert b.kmer_size == 7                                              # 5^7 * 4 B = 312 500 <= 1 MiB < 5^8 * 4 B
