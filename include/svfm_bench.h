/*
 * svfm_bench.h -- measurement and self-check helpers, exported by libsvfm_bench.so (NOT by the product library
 * libsvfm.so, whose ABI is include/svfm.h).  Not part of the reference's interface: synthetic data generation on the device
 * (bench.py builds 1-3 Gbp indexes inside one GPU job), the random-32-byte-sector gather microbenchmark
 * that defines the roofline of SURVEY.md section 8d, and size-independent result checks.
 * All pointers named d_* are device pointers on the current device; everything runs on `stream`
 * (a cudaStream_t, NULL = default stream) and returns an svfm status code.
 */
#ifndef SVFM_BENCH_H
#define SVFM_BENCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Counter-based generator shared with the numpy twin in sview_fmindex_b200/synth.py:
 *   h(seed, i) = splitmix64(seed + 0x9E3779B97F4A7C15 * (i + 1))
 *   text[i]    = alphabet[(h(seed, i) >> 32) * alphabet_len >> 32]
 * `rare` (0 = none) replaces text[i] by rare_byte when h(seed ^ 0xA5A5A5A5, i) % rare == 0
 * (wildcard residues in the protein config). */
int svfm_bench_synth_text(uint8_t* d_text, uint64_t n, uint64_t seed, const uint8_t* alphabet,
                          uint32_t alphabet_len, uint32_t rare, uint8_t rare_byte, void* stream);

/* Pattern i = text[s_i .. s_i+len) with s_i = h(seed, i) % (n - len + 1) (bench/src/generate.rs:105-114
 * draws the start uniformly in 0..=n-len).  d_starts (nullable) receives s_i as u64. */
int svfm_bench_synth_patterns(const uint8_t* d_text, uint64_t n, uint8_t* d_pats, uint64_t* d_starts,
                              uint64_t count, uint32_t len, uint64_t seed, void* stream);

/* Random-sector gather roofline: `loads` independent, uniformly random, 32-byte-aligned 32-byte loads over
 * d_buf[0..bytes) per launch, `iters` timed launches after one warm-up; *sectors_per_s = best launch. */
int svfm_bench_gather32(const uint8_t* d_buf, uint64_t bytes, uint64_t loads, uint32_t iters, uint64_t seed,
                        double* sectors_per_s, double* ms_best, void* stream);

/* Size-independent checks of a locate result (CSR out_offs u64[n+1] + positions, pos_bits 32|64):
 *  violations[0] = reported positions p where text[p..p+len) does not match pattern i symbol-by-symbol
 *                  after encoding through table256 (NULL = identity), or p+len > n
 *  violations[1] = patterns whose own source position d_starts[i] is missing from their list
 *  violations[2] = patterns with an empty list (every pattern was cut from the text)
 *  digest        = sum_i sum_p (p+1)*(2i+1) mod 2^64 (the oracle computes the same number) */
int svfm_bench_verify_locate(const uint8_t* d_text, uint64_t n, const uint8_t* d_pats, uint32_t len,
                             uint64_t n_pats, const uint64_t* d_starts, const uint64_t* d_out_offs,
                             const void* d_positions, uint32_t pos_bits, const uint8_t* table256,
                             uint64_t violations[3], uint64_t* digest, void* stream);

/* sum of counts and order-independent digest of a count result: sum_i count_i * (2i+1) mod 2^64 */
int svfm_bench_count_digest(const void* d_counts, uint32_t pos_bits, uint64_t n, uint64_t* sum, uint64_t* digest,
                            void* stream);

/* Write `bytes` of d_buf (flushes L2 when bytes > L2 size). */
int svfm_bench_flush_l2(uint8_t* d_buf, uint64_t bytes, void* stream);

/* thread-local text of the last failure inside this library */
const char* svfm_bench_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
