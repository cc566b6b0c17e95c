"""Synthetic workloads of BASELINE.json: numpy twin of the device generators in csrc/benchtools.cu
(include/svfm_bench.h), so that small cases can be reproduced on the host bit for bit.

    h(seed, i) = splitmix64(seed + 0x9E3779B97F4A7C15 * (i + 1))
    text[i]    = alphabet[(h(seed, i) >> 32) * len(alphabet) >> 32]
    start_i    = h(seed, i) % (n - len + 1)         (bench/src/generate.rs:105-114: uniform in 0..=n-len)
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_G = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + _G) & _M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return x ^ (x >> np.uint64(31))


def hash_at(seed: int, i: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) + _G * (i.astype(np.uint64) + np.uint64(1)))


def synth_text(n: int, seed: int, alphabet: bytes, rare: int = 0, rare_byte: int = 0) -> np.ndarray:
    i = np.arange(n, dtype=np.uint64)
    h = hash_at(seed, i)
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    text = alpha[((h >> np.uint64(32)) * np.uint64(len(alphabet))) >> np.uint64(32)].copy()
    if rare:
        text[hash_at(seed ^ 0xA5A5A5A5, i) % np.uint64(rare) == 0] = rare_byte
    return text


def synth_pattern_starts(n: int, count: int, length: int, seed: int) -> np.ndarray:
    return hash_at(seed, np.arange(count, dtype=np.uint64)) % np.uint64(n - length + 1)


def synth_patterns(text: np.ndarray, count: int, length: int, seed: int):
    starts = synth_pattern_starts(len(text), count, length, seed)
    idx = starts[:, None].astype(np.int64) + np.arange(length, dtype=np.int64)[None, :]
    return text[idx], starts


NUCLEOTIDES = b"ACGT"
AMINO_ACIDS = b"ACDEFGHIKLMNPQRSTVWY"
