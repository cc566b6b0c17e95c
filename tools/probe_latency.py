"""Single-call latency of the reference-shaped API (one pattern per call) and of small batches, host buffers."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sview_fmindex_b200 import EncodingTable, FmIndex, FmIndexBuilder, IndexType, aligned_empty  # noqa: E402

rng = np.random.default_rng(1)
n = 2_000_000
text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
enc = EncodingTable.from_symbols([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
it = IndexType(32, 3, 64, True)
b = FmIndexBuilder(n, enc.symbol_count(), enc, it)
b.kmer_size, b.sampling_ratio = 3, 2
blob = aligned_empty(b.blob_size())
b.build(text, blob)
ix = FmIndex.load(blob, it)
pats = [bytes(text[s:s + 20]) for s in rng.integers(0, n - 20, size=2000)]
for p in pats[:50]:
    ix.count(p); ix.locate(p)
t0 = time.perf_counter(); [ix.count(p) for p in pats]; t1 = time.perf_counter(); [ix.locate(p) for p in pats]; t2 = time.perf_counter()
print(f"single-pattern calls: count {1e6 * (t1 - t0) / len(pats):.1f} us/call, locate {1e6 * (t2 - t1) / len(pats):.1f} us/call")
for m in (10, 100, 1000, 10000, 100000):
    starts = rng.integers(0, n - 20, size=m)
    batch = text[starts[:, None] + np.arange(20)[None, :]]
    ix.count_batch(batch); ix.locate_batch(batch)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        ix.count_batch(batch)
    t1 = time.perf_counter()
    for _ in range(reps):
        ix.locate_batch(batch)
    t2 = time.perf_counter()
    print(f"batch of {m:6d}: count {1e6 * (t1 - t0) / reps:8.1f} us ({m * reps / (t1 - t0) / 1e6:7.2f} M/s)   "
          f"locate {1e6 * (t2 - t1) / reps:8.1f} us ({m * reps / (t2 - t1) / 1e6:7.2f} M/s)")
