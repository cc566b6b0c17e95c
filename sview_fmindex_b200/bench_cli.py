"""Bench-CLI-compatible front end: the `build` and `locate` subcommands of the reference's bench
binary (`bench/src/main.rs:84-131`) for this engine, on the same data directory layout and file formats, so that rows of
its `run_benchmark.sh` can be reproduced end to end and results diffed line by line:

    python -m sview_fmindex_b200.bench_cli generate-text    -d DIR -t 1000000000 -s 42
    python -m sview_fmindex_b200.bench_cli generate-pattern -d DIR -p 20 -n 100000 -s 42
    python -m sview_fmindex_b200.bench_cli build  -d DIR -s 2 -k 3 [-t]      # text.txt -> sview-memory-block{3,2}.blob
    python -m sview_fmindex_b200.bench_cli locate -d DIR [-t] [-a sview-mmap]  # pattern.txt -> <stem>-results.txt

  * blob file names, symbols (`Aa,Cc,Gg,Tt[,Nn]`, `bench/src/build/mod.rs:28-30`), Block2<u64> with `-t` / Block3<u64>
    without, u32 positions, SASR / KLTS meaning: `bench/src/build/sview_memory.rs:10-107`.  The blob is the reference's
    format byte for byte, so `FmIndex::load` of the Rust crate reads what `build` writes and `locate` reads what the Rust
    `build` wrote.
  * `locate` reads `pattern.txt` (one pattern per line), runs ONE `locate_batch` over all lines instead of one `locate`
    call per line (`bench/src/locate/sview_memory.rs:32-36`), and writes one line per pattern: positions in the reference's
    SA-row order joined by "," (`bench/src/locate/mod.rs:115-123`); absent patterns give an empty line.
  * `locate -a sview-mmap` maps the blob file instead of reading it (`bench/src/locate/sview_mmap.rs:20-46`: `Mmap::map` +
    the optional MMAP_ADVICE_RANDOM / MMAP_ADVICE_SEQUENTIAL / MMAP_ADVICE_DONTDUMP advice, same environment variables) and
    hands the mapping to `FmIndex::load`, which uploads straight out of the page cache; `-a sview-memory` (default) reads
    the file into memory first (`bench/src/locate/sview_memory.rs:19`).
  * `generate-*` use this repo's counter-based generator (sview_fmindex_b200/synth.py), not Rust's `StdRng`, so the bytes
    differ from the reference's for the same seed; shapes and formats are the same (`bench/src/generate.rs:37-45,105-114`).
The timing lines keep the reference's wording ("Blob loading time", "Locate processing time", ...; nanoseconds)."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

SYMBOLS_ACGT = [b"Aa", b"Cc", b"Gg", b"Tt"]           # bench/src/build/mod.rs:29
SYMBOLS_ACGTN = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]   # bench/src/build/mod.rs:30


def blob_stem(treat_t_as_wildcard: bool) -> str:
    return "sview-memory-block2" if treat_t_as_wildcard else "sview-memory-block3"


def index_type(treat_t_as_wildcard: bool):
    from . import IndexType
    return IndexType(32, 2 if treat_t_as_wildcard else 3, 64, True)


def read_patterns(path: str):
    """pattern.txt -> (bytes u8[], offs u64[n+1]): one pattern per line, line terminators stripped
    (BufRead::lines, bench/src/locate/sview_memory.rs:32)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size == 0:
        return raw, np.zeros(1, dtype=np.uint64)
    if raw[-1] != 10:
        raw = np.concatenate([raw, np.array([10], dtype=np.uint8)])
    nl = np.flatnonzero(raw == 10)
    starts = np.concatenate([[0], nl[:-1] + 1])
    ends = nl.copy()
    cr = (ends > starts) & (raw[np.maximum(ends - 1, 0)] == 13)   # "\r\n"
    ends = ends - cr.astype(ends.dtype)
    lens = ends - starts
    keep = np.ones(raw.size, dtype=bool)
    keep[nl] = False
    keep[ends[cr]] = False
    offs = np.zeros(len(lens) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    return np.ascontiguousarray(raw[keep]), offs


def format_results(offs: np.ndarray, pos: np.ndarray) -> bytes:
    """One line per pattern, positions joined by ',' (write_locations_to_file, bench/src/locate/mod.rs:115-123);
    a pattern without occurrences gives an empty line."""
    o = offs.astype(np.int64).tolist()
    p = pos.tolist()
    return "".join(",".join(map(str, p[o[i]:o[i + 1]])) + "\n" for i in range(len(o) - 1)).encode()


def cmd_generate_text(a):
    from . import synth
    os.makedirs(a.data_dir, exist_ok=True)
    path = os.path.join(a.data_dir, "text.txt")
    if os.path.exists(path) and not a.overwrite:
        print(f"Text file already exists: {path}\nUse --overwrite to overwrite.")
        return 0
    t0 = time.perf_counter_ns()
    synth.synth_text(a.text_length, a.seed, synth.NUCLEOTIDES).tofile(path)   # no trailing newline
    print(f"Text file created: {path}")
    print(f"Total time: {time.perf_counter_ns() - t0} ns")
    return 0


def cmd_generate_pattern(a):
    from . import synth
    tpath = os.path.join(a.data_dir, "text.txt")
    ppath = os.path.join(a.data_dir, "pattern.txt")
    if os.path.exists(ppath) and not a.overwrite:
        print(f"Pattern file already exists: {ppath}\nUse --overwrite to overwrite.")
        return 0
    t0 = time.perf_counter_ns()
    text = np.fromfile(tpath, dtype=np.uint8)
    pats, _ = synth.synth_patterns(text, a.pattern_count, a.pattern_length, a.seed)
    lines = np.concatenate([pats, np.full((pats.shape[0], 1), 10, dtype=np.uint8)], axis=1)
    lines.tofile(ppath)
    print(f"Pattern file created: {ppath}")
    print(f"Total time: {time.perf_counter_ns() - t0} ns")
    return 0


def cmd_build(a):
    from . import EncodingTable, FmIndexBuilder, LookupTableConfig, SuffixArrayConfig
    t_total = time.perf_counter_ns()
    print("Building sview-fmindex-b200 (GPU suffix sort, reference blob format)...")
    print(f"SASR: {a.sasr}, KLTS: {a.klts}")
    text = np.fromfile(os.path.join(a.data_dir, "text.txt"), dtype=np.uint8)
    print(f"Loaded text: {text.size} bytes")
    symbols = SYMBOLS_ACGT if a.treat_t_as_wildcard else SYMBOLS_ACGTN
    enc = EncodingTable.from_symbols(symbols)
    b = FmIndexBuilder(text.size, enc.symbol_count(), enc, index_type(a.treat_t_as_wildcard))
    b.set_suffix_array_config(SuffixArrayConfig.Uncompressed if a.sasr == 1 else SuffixArrayConfig.Compressed(a.sasr))
    b.set_lookup_table_config(LookupTableConfig.NONE if a.klts == 1 else LookupTableConfig.KmerSize(a.klts))
    size = b.blob_size()
    print(f"Blob size: {size} bytes")
    from . import aligned_empty
    blob = aligned_empty(size)
    t0 = time.perf_counter_ns()
    b.build(text, blob, device=a.device)
    print(f"Build time: {time.perf_counter_ns() - t0} ns")
    t0 = time.perf_counter_ns()
    out = os.path.join(a.data_dir, blob_stem(a.treat_t_as_wildcard) + ".blob")
    blob.tofile(out)
    print(f"Save time: {time.perf_counter_ns() - t0} ns")
    print(f"Index saved to: {out}")
    print(f"Total time: {time.perf_counter_ns() - t_total} ns")
    return 0


def map_blob(path: str) -> np.ndarray:
    """bench/src/locate/sview_mmap.rs:20-46: read-only mapping of the blob file + the advice the environment asks for."""
    import mmap
    m = np.memmap(path, dtype=np.uint8, mode="r")
    raw = getattr(m, "_mmap", None)
    if raw is not None and hasattr(raw, "madvise"):
        if "MMAP_ADVICE_RANDOM" in os.environ:
            print("Applying MADV_RANDOM advice to mmap")
            raw.madvise(mmap.MADV_RANDOM)
        elif "MMAP_ADVICE_SEQUENTIAL" in os.environ:
            print("Applying MADV_SEQUENTIAL advice to mmap")
            raw.madvise(mmap.MADV_SEQUENTIAL)
        elif "MMAP_ADVICE_DONTDUMP" in os.environ and hasattr(mmap, "MADV_DONTDUMP"):
            print("Applying MADV_DONTDUMP advice to mmap")
            raw.madvise(mmap.MADV_DONTDUMP)
    return m


def cmd_locate(a):
    from . import FmIndex
    t_total = time.perf_counter_ns()
    stem = blob_stem(a.treat_t_as_wildcard)
    blob_path = os.path.join(a.data_dir, stem + ".blob")
    if not os.path.exists(blob_path):
        print(f"{'Block2' if a.treat_t_as_wildcard else 'Block3'} blob file not found: {blob_path}", file=sys.stderr)
        return 1
    print(f"Using blob file: {blob_path}")
    t0 = time.perf_counter_ns()
    if a.algorithm == "sview-mmap":
        blob = map_blob(blob_path)
    else:
        blob = np.fromfile(blob_path, dtype=np.uint8)
    ix = FmIndex.load(blob, index_type(a.treat_t_as_wildcard), device=a.device)
    load_ns = time.perf_counter_ns() - t0
    t0 = time.perf_counter_ns()
    data, offs = read_patterns(os.path.join(a.data_dir, "pattern.txt"))
    n = len(offs) - 1
    lens = np.diff(offs)
    if n and lens.min() == lens.max() and lens[0] > 0:
        out_offs, pos = ix.locate_batch(data.reshape(n, int(lens[0])))      # fixed-length fast path
    else:
        out_offs, pos = ix.locate_batch((data, offs))                         # (bytes, offsets) form
    search_ns = time.perf_counter_ns() - t0
    result_path = os.path.join(a.data_dir, stem + "-results.txt")
    with open(result_path, "wb") as f:
        f.write(format_results(out_offs, pos))
    locate_ns = time.perf_counter_ns() - t0
    print(f"Blob loading time: {load_ns} ns")
    print(f"Locate processing time: {locate_ns} ns")
    print(f"  of which read patterns + search on the GPU: {search_ns} ns ({n} patterns, {int(out_offs[-1])} locations)")
    print(f"Results saved to: {result_path}")
    print(f"Total time: {time.perf_counter_ns() - t_total} ns")
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="sview_fmindex_b200.bench_cli", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="command", required=True)
    g = sub.add_parser("generate-text")
    g.add_argument("-d", "--data-dir", default="test_data")
    g.add_argument("-t", "--text-length", type=int, default=100000)
    g.add_argument("-s", "--seed", type=int, default=0)
    g.add_argument("--overwrite", action="store_true")
    g.set_defaults(fn=cmd_generate_text)
    g = sub.add_parser("generate-pattern")
    g.add_argument("-d", "--data-dir", default="test_data")
    g.add_argument("-p", "--pattern-length", type=int, default=20)
    g.add_argument("-n", "--pattern-count", type=int, default=100)
    g.add_argument("-s", "--seed", type=int, default=0)
    g.add_argument("--overwrite", action="store_true")
    g.set_defaults(fn=cmd_generate_pattern)
    g = sub.add_parser("build")
    g.add_argument("-d", "--data-dir", default="test_data")
    g.add_argument("-s", "--sasr", type=int, default=2)
    g.add_argument("-k", "--klts", type=int, default=3)
    g.add_argument("-t", "--treat-t-as-wildcard", action="store_true")
    g.add_argument("--device", type=int, default=0)
    g.set_defaults(fn=cmd_build)
    g = sub.add_parser("locate")
    g.add_argument("-d", "--data-dir", default="test_data")
    g.add_argument("-t", "--treat-t-as-wildcard", action="store_true")
    g.add_argument("-a", "--algorithm", default="sview-memory", choices=["sview-memory", "sview-mmap"],
                   help="how the blob is loaded (bench/src/locate/mod.rs:44-45)")
    g.add_argument("--device", type=int, default=0)
    g.set_defaults(fn=cmd_locate)
    a = ap.parse_args(argv)
    return a.fn(a)


if __name__ == "__main__":
    sys.exit(main())
