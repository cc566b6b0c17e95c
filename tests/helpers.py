"""Shared test helpers: the independent brute-force matcher and random data in the style of the
reference's tests (sview-fmindex/src/tests/random_data/mod.rs, tests/result_answer/other_crate.rs)."""
import itertools

import numpy as np

ALL_TYPES = [(p, n, v) for p in (32, 64) for n in (2, 3, 4, 5, 6) for v in (32, 64, 128)]


def brute_force_locate(enc_text: np.ndarray, enc_pat: np.ndarray) -> np.ndarray:
    """All start offsets where enc_pat occurs in enc_text (both already symbol-index encoded).
    Independent of any FM-index logic: stands in for the `fm-index 0.1` crate the reference compares
    against (tests/result_answer/other_crate.rs:7-19)."""
    n, m = len(enc_text), len(enc_pat)
    if m == 0 or m > n:
        return np.zeros(0, dtype=np.uint64)
    ok = np.ones(n - m + 1, dtype=bool)
    for j in range(m):
        ok &= enc_text[j:n - m + 1 + j] == enc_pat[j]
    return np.nonzero(ok)[0].astype(np.uint64)


def gen_rand_chr_list(rng, chr_count: int) -> bytes:
    """Distinct printable bytes (tests/random_data/mod.rs:5-16 draws distinct ASCII symbols)."""
    pool = np.arange(33, 127, dtype=np.uint8)
    return bytes(rng.choice(pool, size=chr_count, replace=False))


def gen_rand_text(rng, chr_list: bytes, min_len: int, max_len: int) -> bytes:
    n = int(rng.integers(min_len, max_len + 1))
    return bytes(np.frombuffer(chr_list, dtype=np.uint8)[rng.integers(0, len(chr_list), size=n)])


def gen_rand_pattern(rng, text: bytes, min_len: int, max_len: int) -> bytes:
    ln = int(rng.integers(min_len, min(max_len, len(text)) + 1))
    st = int(rng.integers(0, len(text) - ln + 1))
    return text[st:st + ln]


def type_combos():
    return list(itertools.product((32, 64), (2, 3, 4, 5, 6), (32, 64, 128)))
